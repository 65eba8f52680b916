"""Per-kernel table of key metrics from an .ncu-rep with several kernels (run here, no GPU needed).
usage: python tools/ncu_table.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'),
        ('sm__cycles_elapsed.avg.per_second', 'GHz'),
        ('lts__t_sector_hit_rate.pct', 'l2hit%'),
        ('smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'st_long'),
        ('smsp__average_warp_latency_issue_stalled_lg_throttle.ratio', 'st_lg'),
        ('smsp__average_warp_latency_issue_stalled_membar.ratio', 'st_membar'),
        ('smsp__average_warp_latency_issue_stalled_barrier.ratio', 'st_bar'),
        ('smsp__average_warp_latency_issue_stalled_mio_throttle.ratio', 'st_mio'),
        ('smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio', 'st_short'),
        ]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(name[:110])
    out = []
    for key, short in want:
        if key in hdr:
            i = hdr.index(key)
            out.append(f"{short}={r[i]}{units[i] if short in ('time', 'dram_rd', 'dram_wr') else ''}")
    print("   ", "  ".join(out))
