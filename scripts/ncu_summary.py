"""Print the metrics that matter for this repo's kernels from an ncu report (ncu -i X --page raw --csv)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"

rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("==", d.get("Kernel Name", "")[:110])
    for k in KEYS:
        if k in d:
            print(f"   {k:85s} {d[k]:>16s} {u[k]}")
    st = sorted(((float(v), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                 if k.startswith(STALL) and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
    print("   stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
