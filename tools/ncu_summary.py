"""Key metrics per kernel launch from an .ncu-rep (raw page): usage python tools/ncu_summary.py file.ncu-rep"""
import csv
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]


def dominant_json(path, out_json, match):
    """longest launch whose name contains `match` -> the per-launch numbers bench.py puts into its roofline block"""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = lambda k: hdr.index(k)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dur = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    best = None
    for r in rows[2:]:
        if match not in r[col("Kernel Name")]:
            continue
        ms = float(r[col("gpu__time_duration.sum")]) * dur.get(units[col("gpu__time_duration.sum")], 1.0)
        if best is None or ms > best[0]:
            best = (ms, r)
    if best is None:
        raise SystemExit(f"no launch matching {match}")
    ms, r = best
    val = lambda k: float(r[col(k)]) * scale.get(units[col(k)], 1.0)
    d = {"kernel": r[col("Kernel Name")][:80], "grid": r[col("Grid Size")], "duration_ms_under_ncu": round(ms, 4),
         "dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
         "tensor_pipe_pct": round(float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]), 2),
         "l1tex_throughput_pct": round(float(r[col("l1tex__throughput.avg.pct_of_peak_sustained_elapsed")]), 2),
         "dram_throughput_pct": round(float(r[col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")]), 2),
         "source": "profiles/" + os.path.basename(path).replace(".ncu-rep", "_summary.txt") + " (ncu --set full, longest launch)"}
    with open(out_json, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d))


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
    if len(sys.argv) > 3 and sys.argv[2] == "--json":
        dominant_json(sys.argv[1], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "spconv_tc_kernel")
    else:
        main(sys.argv[1])
