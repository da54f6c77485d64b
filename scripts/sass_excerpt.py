"""Build-container side: trimmed SASS evidence per hot kernel from the built libdctd.so -> profiles/<tag>_sass_<kernel>.txt:
instruction count, opcode histogram, the TMA / mbarrier / SAD / packed-FMA opcodes that prove the sm_100a code path, and
the 70-instruction window richest in the kernel's arithmetic opcode (its inner loop).

    python scripts/sass_excerpt.py r2
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
KERNELS = [
    ('fingerprint', r'fp_ws_kernelILi2ELi1280ELb0E', 'FFMA2', 'fp_ws_plain'),
    ('fingerprint', r'fp_ws_kernelILi2ELi1280ELb1E', 'FFMA2', 'fp_ws_rider'),
    ('l1topk', r'l1_thresh_scan_kernelILi16ELi8ELi2ELi4ELi30E', 'VABSDIFF4', 'l1_thresh_scan'),
    ('l1topk', r'l1_stream_fused_kernelILi8ELi7ELi1ELi30ELi0E', 'VABSDIFF4', 'l1_stream_fused'),
    ('l1topk', r'l1_protein_kernelILi16ELi8ELi2ELi4ELi30E', 'VABSDIFF4', 'l1_protein'),
    ('api', r'gather_kernel', 'LDG', 'gather'),
]
KEY = ['UBLKCP', 'SYNCS', 'VABSDIFF4', 'FFMA2', 'FADD2', 'FMUL2', 'LDS.128', 'LDS.64', 'F2F.F64.F32', 'DADD', 'DFMA', 'REDUX',
       'NANOSLEEP', 'ATOMG', 'ATOM', 'RED', 'BAR.SYNC', 'MEMBAR', 'UTMALDG', 'UTCHMMA', 'HMMA']

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'dctdomain_b200', 'libdctd.so')], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cache = {}
for obj, pat, arith, name in KERNELS:
    if obj not in cache:
        cubin = [f for f in os.listdir(tmp) if f.startswith(obj) and f.endswith('.cubin')][0]
        cache[obj] = subprocess.run(['nvdisasm', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    ins, inside, title = [], False, ''
    for ln in cache[obj].splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            inside = re.search(pat, ln) is not None
            if inside:
                title = ln.strip('/- ').replace('.text.', '')
            continue
        if inside:
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
            if m:
                ins.append((m.group(1), m.group(2).strip()))
    ops = Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0] for _, t in ins)
    best, best_i = -1, 0
    hits = [1 if arith in t else 0 for _, t in ins]
    W = 70
    run = sum(hits[:W])
    for i in range(0, max(1, len(ins) - W)):
        if run > best:
            best, best_i = run, i
        run += (hits[i + W] if i + W < len(hits) else 0) - hits[i]
    with open(os.path.join(ROOT, 'profiles', f'{tag}_sass_{name}.txt'), 'w') as f:
        f.write(f'{title}\nlibdctd.so built with nvcc -gencode arch=compute_100a,code=sm_100a; nvdisasm -c of the embedded cubin\n')
        f.write(f'instructions: {len(ins)}\n\nkey opcodes (count in the kernel):\n')
        for k in KEY:
            c = sum(v for o, v in ops.items() if o.startswith(k))
            if c:
                f.write(f'  {k:14s} {c}\n')
        f.write('\nopcode histogram (top 30):\n')
        for o, c in ops.most_common(30):
            f.write(f'  {o:34s} {c}\n')
        f.write(f'\n{W}-instruction window richest in {arith} ({best} of them), from /*{ins[best_i][0]}*/:\n')
        for a, t in ins[best_i:best_i + W]:
            f.write(f'  /*{a}*/ {t}\n')
    print(name, len(ins), {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in ('UBLKCP', 'SYNCS', 'VABSDIFF4', 'FFMA2')})
