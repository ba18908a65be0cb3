"""Key metrics per kernel launch from an .ncu-rep (raw page): usage python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    stall = [i for i, h in enumerate(hdr) if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                print(f"   {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        st = sorted(((float(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:8]
        for v, h in st:
            print(f"   stall {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):40s} {v:8.2f} %")


if __name__ == "__main__":
    main(sys.argv[1])
