"""Summarise an .ncu-rep (raw + source pages) the way profiles/*.md quote it.
usage: python profiles/ncu_summarize.py gpurun_out/prof.ncu-rep [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            print("%-70s %s %s" % (k, r[hdr.index(k)][:90], units[hdr.index(k)]))
    print("-" * 40)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith('0x')]
# the report repeats the listing per captured launch: keep the first
first = body[0][0]
for i in range(1, len(body)):
    if body[i][0] == first:
        body = body[:i]
        break
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(int(r[ix[k]] or 0) for r in body) for k in stalls}
print("samples", tot, {k[6:]: "%.1f%%" % (100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > tot * 0.01})
cum = 0
for i, r in enumerate(body):
    s = int(r[ix['# Samples']] or 0)
    cum += s
    ins = r[1].strip()
    if s >= minpct / 100 * tot or any(k in ins for k in ('BAR.SYNC', 'SYNCS.PHASECHK', 'UBLKCP', 'EXIT', 'MUFU.RCP64H')):
        print('%4d cum=%5.1f%% %5.1f%% ex=%9s thr=%4s %-60s %s' % (i, 100 * cum / tot, 100 * s / tot, r[ix['Instructions Executed']], r[ix['Avg. Predicated-On Threads Executed']], ins[:60],
              ' '.join('%s=%s' % (k[6:], r[ix[k]]) for k in stalls if int(r[ix[k]] or 0) > 0.25 * max(s, 1) and s >= minpct / 100 * tot)))
