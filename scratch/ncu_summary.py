"""Dump a compact text summary of an .ncu-rep (key raw metrics + top stall sites) for profiles/."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__cluster_max_active', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_fmaheavy.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('kernel:', r[h.index('Kernel Name')][:80])
    for i, n in enumerate(h):
        if n in want:
            print('  %-70s %-14s %s' % (n, u[i], r[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; idx = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h) and r != h and r[idx['# Samples']].strip().isdigit()]   # (several results: later header rows are dropped, the sites of all results are pooled)
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print('SASS instructions: %d, warp stall samples: %d' % (len(data), tot))
agg = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        agg[s] += int(r[idx[s]] or 0)
print('stall reasons: ' + ', '.join('%s %.1f%%' % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
ops = {}
for r in data:
    t = r[idx['Source']].split()
    if not t: continue
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] = ops.get(op, 0) + int(r[idx['Instructions Executed']] or 0)
ex = sum(ops.values())
print('executed instruction mix: ' + ' '.join('%s:%.1f%%' % (k, 100.0 * v / max(ex, 1)) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:14]))
print('top stall sites:')
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:12]:
    st = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print('  %5s samples  x%-8s %-70s %s' % (r[idx['# Samples']], r[idx['Instructions Executed']], r[idx['Source']][:70], st))
