"""Key metrics of every launch in an .ncu-rep (ncu --set full capture), one block per launch:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_ncu_x.txt"""
import csv, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
print("# %s: %d launches (ncu --set full --clock-control none; per-launch, cold cache)" % (sys.argv[1].split("/")[-1], len(rows) - 2))
for r in rows[2:]:
    d = dict(zip(h, r))
    print("\n== %s  grid %s block %s" % (d.get("Kernel Name"), d.get("launch__grid_size"), d.get("launch__block_size")))
    for i, c in enumerate(h):
        if c in WANT:
            print("  %-82s %14s %s" % (c, r[i], units[i]))
    try:
        u = dict(zip(h, units))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        t = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]] + float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
        print("  %-82s %14.1f MB" % ("traffic = dram read + write", t / 1e6))
    except Exception:
        pass
