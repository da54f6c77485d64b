"""Build-container side: turn gpurun_out/{launches,fp,l1}_<tag> into tracked summaries under profiles/.

    python scripts/summarize_profile.py r1c
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
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


summary = {'tag': tag, 'command': 'python bench.py --steps 4 --warmup 3 --no-cpu (scripts/profile.sh)'}
# launch list
lp = os.path.join(G, f'launches_{tag}.csv')
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        val = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ns = val * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
        name = r['Kernel Name'].split('(')[0][-70:]
        tot[name][0] += 1
        tot[name][1] += ns
    total = sum(v[1] for v in tot.values())
    summary['launch_list'] = [{'kernel': k, 'launches': v[0], 'total_ms': round(v[1] / 1e6, 3), 'share': round(v[1] / total, 4)}
                              for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]]
    with open(os.path.join(P, f'{tag}_launches.csv'), 'w') as f:
        f.writelines(lines)
for name in ('fp', 'l1'):
    rep = os.path.join(G, f'{name}_{tag}.ncu-rep')
    if os.path.exists(rep):
        summary[name + '_kernel'] = raw(rep)
json.dump(summary, open(os.path.join(P, f'{tag}_summary.json'), 'w'), indent=1)

# roofline traffic (per launch) for bench.py
traffic = {}
tp = os.path.join(P, 'roofline_traffic.json')
if os.path.exists(tp):
    traffic = json.load(open(tp))


def to_bytes(s):
    v, u = s.split()
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u]


for name, key in (('fp', 'fp_ws_kernel'), ('l1', 'l1_scan_kernel')):
    ks = summary.get(name + '_kernel')
    if ks:
        vals = [to_bytes(k['dram__bytes_read.sum']) + to_bytes(k['dram__bytes_write.sum']) for k in ks]
        traffic[key] = sum(vals) / len(vals)
        traffic[key + '_source'] = f'profiles/{tag}_summary.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)'
json.dump(traffic, open(tp, 'w'), indent=1)
print(json.dumps(summary, indent=1)[:6000])
