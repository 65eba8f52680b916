"""Key metrics of one kernel from an .ncu-rep (run here, no GPU needed).
usage: python tools/ncu_summary.py file.ncu-rep [src-threshold]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.avg.per_second',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"  {h} [{u}] = {v}")
