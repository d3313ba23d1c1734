"""TEST INFRASTRUCTURE ONLY: CPU oracle of the TAI hot path (see oracle/oracle.py)."""
