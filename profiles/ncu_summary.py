"""One line per kernel of an `ncu --set full` report: duration, tensor-pipe activity, DRAM bytes, grid, registers.
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]      (needs `ncu` on PATH; reads, never profiles)"""
import csv
import io
import subprocess
import sys

WANT = {"gpu__time_duration.sum": "gpu__time_duration_us", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active":
        "sm__pipe_tensor_cycles_active%", "dram__bytes_read.sum": "dram__bytes_read_MB", "dram__bytes_write.sum":
        "dram__bytes_write_MB", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "gpu__dram_throughput%",
        "lts__t_bytes.sum": "l2_bytes_MB", "launch__grid_size": "launch__grid_size", "launch__registers_per_thread":
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm__throughput%",
        "launch__occupancy_limit_shared_mem": "occupancy_limit_smem", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct":
        "stall_long_scoreboard%", "smsp__warp_issue_stalled_barrier_per_warp_active.pct": "stall_barrier%",
        "sm__cycles_active.avg": "sm__cycles_active", "sm__cycles_elapsed.avg": "sm__cycles_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active%", "lts__t_sector_hit_rate.pct": "l2_hit_rate%",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput%",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_mem_cycles_active%",
        "launch__shared_mem_per_block_dynamic": "smem_dynamic_KB", "launch__occupancy_limit_registers": "occupancy_limit_regs",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct": "stall_membar%",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct": "stall_short_scoreboard%",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct": "stall_wait%",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct": "stall_sleeping%",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct": "stall_math_throttle%",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct": "stall_lg_throttle%"}

for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    print("# %s" % rep)
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        d = {"Kernel Name": r[names.index("Kernel Name")][:64]}
        for col, label in WANT.items():
            if col in names:
                i = names.index(col)
                v, u = r[i].replace(",", ""), units[i]
                try:
                    x = float(v)
                    if u in ("byte", "bytes"):
                        x /= 1e6
                    elif u == "Kbyte":
                        x /= 1e3
                    elif u == "Gbyte":
                        x *= 1e3
                    elif u in ("nsecond", "ns"):
                        x /= 1e3
                    elif u in ("msecond", "ms"):
                        x *= 1e3
                    d[label] = round(x, 3)
                except ValueError:
                    d[label] = v
        print(d)
