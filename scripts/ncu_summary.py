"""Summarise an .ncu-rep: key raw metrics per kernel + top stall lines from the source page.
    python scripts/ncu_summary.py <rep> [top_n] [traffic_key]
With `traffic_key` the measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the first kernel in the
report is recorded in profiles/ncu_traffic.json under that key -- bench.py reads `roofline.traffic` from there."""
import csv, json, os, subprocess, sys, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
traffic_key = sys.argv[3] if len(sys.argv) > 3 else None
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
def _bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
if traffic_key and len(rows) > 2:
    r = rows[2]
    tot = _bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + \
          _bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[traffic_key] = {'dram_bytes_per_launch': tot, 'grid_size': int(float(r[idx['launch__grid_size']].replace(',', ''))),
                      'kernel': r[idx['Kernel Name']][:90], 'report': os.path.basename(rep),
                      'duration_us_under_ncu': r[idx['gpu__time_duration.sum']] + ' ' + units[idx['gpu__time_duration.sum']]}
    json.dump(d, open(path, 'w'), indent=1)
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
