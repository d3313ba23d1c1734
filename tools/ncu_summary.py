"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + stall reasons + hottest SASS."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path, topn=14):
    raw = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print("==", d.get("Kernel Name", "?")[:110], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print("  %-78s %s %s" % (k, d[k], units[hdr.index(k)]))
        stalls = sorted(((float(v), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and v),
                        reverse=True)
        print("  stalls per issue:", ", ".join("%s=%.2f" % (k.split("stalled_")[1].split("_per_")[0], v) for v, k in stalls[:9]))
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    h = next(r for r in src if "Address" in r)
    ia, isrc, isamp = h.index("Address"), h.index("Source"), h.index("# Samples")
    data = [r for r in src if len(r) == len(h) and r[ia].startswith("0x")]
    seen, first = set(), []
    for r in data:
        if r[ia] in seen:
            break
        seen.add(r[ia])
        first.append(r)
    tot = sum(int(r[isamp]) for r in first) or 1
    base = int(first[0][ia], 16)
    print("  hottest SASS (of %d samples):" % tot)
    for r in sorted(first, key=lambda r: -int(r[isamp]))[:topn]:
        print("    %6s %5.1f%%  %s" % (hex(int(r[ia], 16) - base), 100.0 * int(r[isamp]) / tot, r[isrc].strip()[:100]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
