"""Build-container side: per-source-line totals of an ncu capture (taken with --import-source on, code built with
-lineinfo).  Joins the SASS page of the report with nvdisasm's line table of the same kernel (instruction order).

    python scripts/ncu_lines.py <report.ncu-rep> <kernel mangled-name regex> [object=fingerprint] [top=40]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(obj, pattern, want_len=None):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'dctdomain_b200', 'libdctd.so')], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith(obj) and f.endswith('.cubin')][0]
    text = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    funcs, out, inside, cur = [], None, False, ('?', 0)
    for ln in text.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            inside = re.search(pattern, ln) is not None
            if inside:
                out = []
                funcs.append(out)
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    # several instantiations may match the pattern: take the one with as many instructions as the report holds
    if want_len is not None:
        for f in funcs:
            if len(f) == want_len:
                return f
    return funcs[0] if funcs else []


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    obj = sys.argv[3] if len(sys.argv) > 3 else 'fingerprint'
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    csv_text = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_text.splitlines()))
    hdr, data = rows[1], rows[2:]
    i_s, i_n = hdr.index('# Samples'), hdr.index('Instructions Executed')
    i_w, i_wi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
    stall = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
    table = line_table(obj, pattern, len(data))
    if len(table) != len(data):
        print(f'warning: {len(table)} instructions in the cubin vs {len(data)} in the report', file=sys.stderr)
    agg = defaultdict(lambda: defaultdict(float))
    tot_s = tot_n = 0
    for (off, loc, ins), r in zip(table, data):
        a = agg[loc]
        s, n = float(r[i_s] or 0), float(r[i_n] or 0)
        a['samples'] += s
        a['inst'] += n
        a['wave'] += float(r[i_w] or 0)
        a['wave_ideal'] += float(r[i_wi] or 0)
        for h, i in stall.items():
            a[h] += float(r[i] or 0)
        tot_s += s
        tot_n += n
    print(f'total samples {tot_s:.0f}, warp instructions {tot_n:.0f}')
    print(f'{"file:line":28s} {"samples%":>8s} {"inst%":>7s} {"smem wf":>11s} {"ideal":>11s}  top stalls')
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        st = sorted(((v, h) for h, v in a.items() if h.startswith('stall_') and v > 0), reverse=True)[:3]
        sts = ', '.join(f'{h[6:]} {v / max(a["samples"], 1):.0%}' for v, h in st)
        print(f'{loc[0] + ":" + str(loc[1]):28s} {a["samples"] / tot_s:8.2%} {a["inst"] / tot_n:7.2%} {a["wave"]:11.0f} {a["wave_ideal"]:11.0f}  {sts}')


if __name__ == '__main__':
    main()
