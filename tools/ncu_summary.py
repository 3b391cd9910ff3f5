"""Key metrics of an .ncu-rep (one kernel): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__inst_executed_pipe_tensor', 'sm__pipe_tensor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput', 'lts__throughput', 'dram__throughput', 'l1tex__throughput',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared',
        'smem', 'tma', 'tensor', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_red']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
pat = sys.argv[2:] or KEYS
for h, u, v in zip(hdr, units, vals):
    if any(k in h for k in pat):
        print('%-80s %-12s %s' % (h, u, v))
