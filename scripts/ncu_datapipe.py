"""Shared-memory / L1 data-pipe accounting of an .ncu-rep (--set full): tcgen05 operand fetch vs LSU wavefronts vs cycles.
    python scripts/ncu_datapipe.py <rep>"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
def g(r, k):
    try: return float(r[h.index(k)].replace(',', ''))
    except Exception: return float('nan')
seen = set()
for r in rows[2:]:
    name = r[h.index('Kernel Name')][:60]
    if name in seen: continue
    seen.add(name)
    sms = 148.0
    cyc = g(r, 'sm__cycles_elapsed.avg')
    tc = g(r, 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum') / sms
    lsu = g(r, 'l1tex__data_pipe_lsu_wavefronts.sum') / sms if 'l1tex__data_pipe_lsu_wavefronts.sum' in h else float('nan')
    lsu_sh = g(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / sms
    print("%-60s %7.1f us  cycles %8.0f  tensor %5.1f %%  | data pipe: tc %5.1f %%  lsu %5.1f %% (shared %5.1f %%)  total %5.1f %%" % (
        name, g(r, 'gpu__time_duration.sum'), cyc, g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        100 * tc / cyc, g(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'), 100 * lsu_sh / cyc,
        100 * tc / cyc + g(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')))
