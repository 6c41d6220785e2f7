"""Prints the key metrics of every kernel in an .ncu-rep (read on the CPU box: ncu -i ... --page raw --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
extra = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:70], 'id', r[0])
    for k in KEYS:
        if k in hdr:
            print(f'   {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
    stalls = sorted(((float(r[hdr.index(k)] or 0), k) for k in extra), reverse=True)[:7]
    for v, k in stalls:
        print(f'   stall {k.split("issue_stalled_")[1].split("_per_warp")[0]:40s} {v:8.2f} %')
