"""Summarise an .ncu-rep: key raw metrics per kernel + top stall lines from the source page."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed_pipe_xu.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']][:70]
    if name in seen: continue
    seen.add(name)
    print('==', name)
    for k in keys:
        if k in idx: print('   %-62s %s %s' % (k, r[idx[k]], units[idx[k]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
sections = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; sections.append(cur)
    elif cur is not None: cur['rows'].append(r)
seen = set()
for s in sections:
    nm = s['name'][:70]
    if nm in seen or not s['rows']: continue
    seen.add(nm)
    h = s['rows'][0]; ix = {x: i for i, x in enumerate(h)}
    data = [r for r in s['rows'][1:] if len(r) > ix['# Samples'] and r[ix['# Samples']].isdigit()]
    tot = sum(int(r[ix['# Samples']]) for r in data)
    print('==== stalls:', nm, 'samples', tot)
    stall_cols = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    agg = collections.Counter()
    for r in data:
        for c in stall_cols:
            v = r[ix[c]]
            if v.isdigit(): agg[c] += int(v)
    print('   by reason:', [(k, v) for k, v in agg.most_common(7)])
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:topn]:
        st = sorted([(int(r[ix[c]]), c) for c in stall_cols if r[ix[c]].isdigit()], reverse=True)[:2]
        print('   %6s %-64s %s' % (r[ix['# Samples']], r[ix['Source']][:64], st))
