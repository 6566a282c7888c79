"""Summarise an ncu report: key raw metrics + opcode histogram per chain step (dev tool)."""
import collections
import csv
import subprocess
import sys

rep, steps = sys.argv[1], float(sys.argv[2])
out = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
d = dict(zip(rows[0], rows[-1]))
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.per_cycle_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active']
keys += [k for k in d if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
for k in keys:
    print(k, '=', d.get(k))
print('warp inst per chain step', float(d['smsp__inst_executed.sum']) / steps)
out = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
hist, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= ia or not r[ia].isdigit():
        continue
    t = r[isrc].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    hist[op] += int(r[ia])
    samp[op] += int(r[isamp])
tot = sum(samp.values())
for op, c in hist.most_common(28):
    print(f'{op:10s} {c/steps:7.1f}/step  samples {100*samp[op]/tot:5.1f}%')
