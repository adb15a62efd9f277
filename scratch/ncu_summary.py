"""Summarise an .ncu-rep (raw page) per kernel launch: python scratch/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__waves_per_multiprocessor',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    tens = [h for h in hdr if ('pipe_tensor' in h or 'tmem' in h.lower() or 'umma' in h.lower() or 'utc' in h.lower()) and '.avg' in h and 'realtime' not in h][:10]
    stall = [h for h in hdr if 'warps_issue_stalled' in h and h.endswith('ratio')]
    print(f"# {path}")
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print(f"===== {name[:90]}")
        for w in WANT + [t for t in tens if t not in WANT]:
            if w in hdr and r[hdr.index(w)] not in ('', 'n/a'):
                print(f"   {w} [{units[hdr.index(w)]}] = {r[hdr.index(w)]}")
        st = sorted([(float(r[hdr.index(h)] or 0), h) for h in stall], reverse=True)[:7]
        print("   stalls per issue: " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per')[0]} {v:.2f}" for v, h in st))
