#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/ (runs where ncu is installed; no GPU needed to read a report):

    python tools/ncu_summary.py gpurun_out/frame_r2.ncu-rep profiles/r2_ncu_full_frame_kernels.txt "header line ..."
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
           "launch__cluster_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def main():
    rep, out, header = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    H, units = rows[0], rows[1]
    lines = ["# ncu --set full --clock-control none --import-source on, one launch of every hot kernel", "# " + header, ""]
    # one launch per kernel name: the longest one (a report may hold many launches of one kernel at different sizes)
    best = {}
    ti = H.index("gpu__time_duration.sum")
    for r in rows[2:]:
        name = r[H.index("Kernel Name")]
        t = float(r[ti].replace(",", ""))
        if name not in best or t > best[name][0]:
            best[name] = (t, r)
    for name, (_, r) in best.items():
        lines.append("== " + name[:110])
        for m in METRICS:
            if m in H:
                lines.append(f"   {m} [{units[H.index(m)]}] = {r[H.index(m)]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
