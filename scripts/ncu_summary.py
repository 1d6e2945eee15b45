"""Summarises an .ncu-rep: key raw metrics per kernel + SASS opcode histogram + top shared-memory conflict sites."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('=' * 100)
    for w in want:
        if w in idx:
            print(f"{w:72s} {r[idx[w]][:90]:>24s} {units[idx[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kernels, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        kernels.append(cur)
        continue
    if cur is None:
        continue
    if cur['hdr'] is None:
        cur['hdr'] = r
        continue
    cur['rows'].append(r)
for k in kernels:
    h = {n: i for i, n in enumerate(k['hdr'])}
    print('-' * 100)
    print(k['name'][:110], '|', len(k['rows']), 'SASS instructions')
    ops, stall, tot, samples = collections.Counter(), collections.Counter(), 0, 0
    for r in k['rows']:
        ins = r[h['Source']].strip()
        n = int(r[h['Instructions Executed']])
        s = int(r[h['# Samples']])
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ins)
        base = (m.group(2) if m else ins).split('.')[0]
        ops[base] += n
        stall[base] += s
        tot += n
        samples += s
    print("total warp instructions", tot, "| stall samples", samples)
    for op, n in ops.most_common(22):
        print(f"  {op:10s} {n:12d} {100*n/max(tot,1):5.1f}%   stall-samples {100*stall[op]/max(samples,1):5.1f}%")
    col = 'L1 Wavefronts Shared Excessive'
    if col in h:
        confl = sorted(((int(r[h[col]]), r[h['Source']].strip(), int(r[h['Instructions Executed']])) for r in k['rows']
                        if r[h[col]] not in ('', '0')), reverse=True)
        print("  top shared-memory conflict sites (excess wavefronts, SASS, executions):")
        for c in confl[:8]:
            print("   ", c)
    top = sorted(((int(r[h['# Samples']]), r[h['Source']].strip()) for r in k['rows']), reverse=True)[:14]
    print("  top stall-sample instructions:")
    for t in top:
        print("   ", t)
