"""Key figures of every kernel in an ncu report: python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [('time us', 'gpu__time_duration.sum'), ('regs', 'launch__registers_per_thread'), ('grid', 'launch__grid_size'),
        ('occ_lim_regs', 'launch__occupancy_limit_registers'), ('occ_lim_smem', 'launch__occupancy_limit_shared_mem'),
        ('warps_active%', 'sm__warps_active.avg.pct_of_peak_sustained_active'), ('inst', 'smsp__inst_executed.sum'),
        ('issue%', 'sm__inst_issued.avg.pct_of_peak_sustained_active'), ('l1tex%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('lts%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), ('dram%', 'dram__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('dram_rd MB', 'dram__bytes_read.sum'), ('dram_wr MB', 'dram__bytes_write.sum'),
        ('fma%', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'), ('alu%', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'),
        ('lsu_wavefronts', 'l1tex__data_pipe_lsu_wavefronts.sum'), ('bank_conf', 'l1tex__data_bank_conflicts_pipe_lsu.sum'),
        ('L1 hit%', 'l1tex__t_sector_hit_rate.pct'), ('L2 hit%', 'lts__t_sector_hit_rate.pct')]
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
seen = set()
for r in data:
    name = r[idx['Kernel Name']].split('(')[0]
    if name in seen and '--all' not in sys.argv:
        continue
    seen.add(name)
    print('==', name)
    print('   ' + '  '.join(f"{k}={r[idx[m]]}{'' if units[idx[m]] in ('', '%') else ' ' + units[idx[m]]}" for k, m in want if m in idx))
    st = sorted(((float(r[idx[h]]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for h in stalls), reverse=True)
    print('   stalls: ' + '  '.join(f"{n}={v:.2f}" for v, n in st[:7]))
