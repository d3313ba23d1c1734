/* TEST INFRASTRUCTURE ONLY -- see THC.h in this directory. */
#include "THC.h"
