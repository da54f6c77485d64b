"""Build-container side: gpurun_out/{launches,<kernel>_raw.csv,<kernel>_lines.txt}_<tag> (written on the GPU box by
scripts/profile_r2.sh) -> tracked summaries under profiles/.     python scripts/summarize_r2.py r2a"""
import csv
import json
import os
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for k in KEEP:
            if k in hdr:
                d[k] = f'{r[hdr.index(k)]} {units[hdr.index(k)]}'.strip()
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    if float(r[i]) >= 0.05:
                        d['stall_' + h.split('stalled_')[1].split('_per')[0]] = round(float(r[i]), 3)
                except ValueError:
                    pass
        res.append(d)
    return res


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    tot = defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        val = float(r['Metric Value'].replace(',', ''))
        ns = val * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(r['Metric Unit'], 1)
        name = r['Kernel Name'].split('(')[0][-70:]
        tot[name][0] += 1
        tot[name][1] += ns
    total = sum(v[1] for v in tot.values())
    return lines, [{'kernel': k, 'launches': v[0], 'total_ms': round(v[1] / 1e6, 3), 'avg_ms': round(v[1] / v[0] / 1e6, 4),
                    'share': round(v[1] / total, 4)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]]


summary = {'tag': tag, 'script': 'scripts/profile_r2.sh (every ncu run follows a plain run of the same command)'}
for name, cmd in (('launches', 'python bench.py --steps 4 --warmup 3 --no-cpu --no-cfg4 --no-allvsall --no-dctsim'),
                  ('launches_stream', 'python scripts/l1_phases.py stream,ref13')):
    path = os.path.join(G, f'{name}_{tag}.csv')
    if os.path.exists(path):
        lines, top = launches(path)
        summary[name] = {'command': cmd, 'per_kernel': top}
        with open(os.path.join(P, f'{tag}_{name}.csv'), 'w') as f:
            f.writelines(lines)
for name, cmd in (('fp', 'bench command, fp_ws_kernel<2,1280,0> launch 5'), ('fprider', 'python scripts/fp_protein.py 2048 4'),
                  ('l1stream', 'python scripts/l1_phases.py stream (8 queries x 6.25M)'),
                  ('l1prot', 'python scripts/dctsim_run.py 4000')):
    path = os.path.join(G, f'{name}_{tag}_raw.csv')
    if os.path.exists(path):
        summary[name] = {'command': cmd, 'ncu': '--set full --clock-control none --import-source on', 'kernels': raw(path)}
        lp = os.path.join(G, f'{name}_{tag}_lines.txt')
        if os.path.exists(lp):
            shutil.copy(lp, os.path.join(P, f'{tag}_{name}_lines.txt'))
json.dump(summary, open(os.path.join(P, f'{tag}_summary.json'), 'w'), indent=1)
print(json.dumps(summary, indent=1)[:9000])
