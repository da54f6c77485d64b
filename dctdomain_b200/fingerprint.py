"""Drop-in for the reference ``src/fingerprint.py`` with ``quantize`` running on the B200.

Same class, field and method names as the reference ``Fingerprint`` (src/fingerprint.py:17-201):
``quantize(qdim)`` fills ``self.quants`` (domain string -> int array of n*m values per layer,
layer order = ``embed`` insertion order) and rewrites ``self.domains``, exactly what
``make_db.queue_cpu`` (src/make_db.py:19-33) and ``Database.add_fprint`` (src/database.py:207)
consume.  The arithmetic runs in libdctd's CUDA kernel (csrc/fingerprint.cu) - there is no CPU
path: without a CUDA device or the built library every entry point raises.

Additions over the reference surface:
  * ``quantize_batch(fps, qdim)``  - one kernel launch for a list of proteins, the replacement
    for ``fprint_cpu``'s ``Pool.starmap(queue_cpu, ...)`` (src/make_db.py:48-49);
  * windowed input: ``embed[layer]`` may be a *list* of the maxlen windows that
    ``Embedding.embed_seq`` (src/embedding.py:153-192) would stitch; the 200-row overlap average
    is applied inside the kernel while the rows are loaded;
  * ``embed`` values may be numpy arrays (host), torch CPU tensors (pinned or not) or torch
    CUDA tensors (consumed in place, no copy).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess as sp
from dataclasses import dataclass, field
from operator import itemgetter

import numpy as np
import torch

from . import _lib

OVERLAP = 200          # src/embedding.py:163
DEFAULT_MAXLEN = 500   # src/make_db.py --maxlen default

_workspaces: dict = {}


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('dctdomain_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(device)


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    ws = _workspaces.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[device] = ws
    return ws


def parse_domain(dom: str, n_rows: int):
    """Rows of a RecCut domain string, with get_doms' bound handling (src/fingerprint.py:160-171).

    Returns ([(begin0, end_exclusive), ...], kept_string).  A segment is dropped when its begin
    lies beyond the protein, an end beyond the protein is clipped, and - as in the reference,
    which removes from the list it iterates over - the segment after a dropped one is passed
    over but stays in the returned string.
    """
    parts = dom.split(',')
    segs = []
    pos = 0
    while pos < len(parts):
        beg_s, end_s = parts[pos].split('-')
        beg, end = int(beg_s), int(end_s)
        if (beg or end) > n_rows:
            parts.pop(pos)
            pos += 1
            continue
        lo = beg - 1 if beg >= 1 else n_rows + beg - 1
        hi = min(end, n_rows)
        segs.append((lo, max(lo, hi)))
        pos += 1
    return segs, ','.join(parts)


class _Plan:
    """Owns a dctd_fp_plan handle."""

    def __init__(self, geo: _lib.FpGeometry, keep):
        self._keep = keep
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().dctd_fp_plan_create(C.byref(geo), C.byref(self.handle)), 'dctd_fp_plan_create')
        L = _lib.lib()
        self.workspace_bytes = int(L.dctd_fp_workspace_bytes(self.handle))
        self.algorithmic_bytes = int(L.dctd_fp_algorithmic_bytes(self.handle))
        self.n_items = int(L.dctd_fp_num_items(self.handle))

    def __del__(self):
        if getattr(self, 'handle', None) and self.handle.value:
            _lib.lib().dctd_fp_plan_destroy(self.handle)
            self.handle = C.c_void_p()


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def make_plan(n_layers, D, n, m, src_rows, prot_src0, prot_nsrc, dom_prot, dom_seg_off, seg_beg, seg_end,
              maxlen=DEFAULT_MAXLEN, overlap=OVERLAP) -> _Plan:
    """Host-side work decomposition for one batch (dctd_fp_plan_create)."""
    arrs = [_i32(x) for x in (src_rows, prot_src0, prot_nsrc, dom_prot, dom_seg_off, seg_beg, seg_end)]
    geo = _lib.FpGeometry()
    geo.n_layers, geo.D, geo.n, geo.m = int(n_layers), int(D), int(n), int(m)
    geo.maxlen, geo.overlap = int(maxlen), int(overlap)
    geo.n_src, geo.src_rows = len(arrs[0]), arrs[0].ctypes.data
    geo.n_prot, geo.prot_src0, geo.prot_nsrc = len(arrs[1]), arrs[1].ctypes.data, arrs[2].ctypes.data
    geo.n_dom, geo.dom_prot, geo.dom_seg_off = len(arrs[3]), arrs[3].ctypes.data, arrs[4].ctypes.data
    geo.seg_beg, geo.seg_end = arrs[5].ctypes.data, arrs[6].ctypes.data
    return _Plan(geo, arrs)


def execute_plan(plan: _Plan, src_tensors, out: torch.Tensor, tables_resident=False, workspace=None):
    """Launch the fingerprint kernel.  ``src_tensors[layer][s]`` are CUDA float32 [rows, D] tensors."""
    dev = out.device
    ptrs = np.array([t.data_ptr() for layer in src_tensors for t in layer], dtype=np.uint64)
    ld = src_tensors[0][0].stride(0) if src_tensors and src_tensors[0] else 0
    ws = workspace if workspace is not None else _workspace(dev, plan.workspace_bytes)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = _lib.lib().dctd_fp_execute(plan.handle, ptrs.ctypes.data, ld, out.data_ptr(), out.stride(0),
                                        ws.data_ptr(), ws.numel(),
                                        _lib.FP_TABLES_RESIDENT if tables_resident else 0, stream)
    _lib.check(rc, 'dctd_fp_execute')
    return out


def _as_cuda(x, device):
    """numpy / torch (cpu, pinned, cuda) [rows, D] -> contiguous float32 CUDA tensor."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def quantize_batch(fps, qdim=(3, 80, 3, 80), device=None, maxlen=DEFAULT_MAXLEN, overlap=OVERLAP):
    """``quantize`` for a list of Fingerprint-like objects in one kernel launch per (n, m) group.

    Each object needs ``embed`` ({layer: [L, D] array | list of window arrays}), ``domains`` (list of
    RecCut strings) and ``quants`` (dict); they are updated exactly as the reference ``quantize`` does
    (src/fingerprint.py:184-201).  Returns the list for convenience.
    """
    fps = list(fps)
    if not fps:
        return fps
    dev = _device(device)
    n_layers = len(fps[0].embed)
    qdim = list(qdim)
    if len(qdim) < 2 * n_layers:
        raise IndexError('qdim needs an (n, m) pair per embedding layer')  # reference: list index error
    for fp in fps:
        if len(fp.embed) != n_layers:
            raise ValueError('all proteins of a batch must carry the same layers')

    # ---- sources (device residency) and geometry ----
    src = [[] for _ in range(n_layers)]
    src_rows, prot_src0, prot_nsrc, prot_len = [], [], [], []
    D = None
    for fp in fps:
        layers = list(fp.embed.values())
        wins0 = layers[0] if isinstance(layers[0], (list, tuple)) else [layers[0]]
        prot_src0.append(len(src_rows))
        prot_nsrc.append(len(wins0))
        for li, lay in enumerate(layers):
            wins = lay if isinstance(lay, (list, tuple)) else [lay]
            if len(wins) != len(wins0):
                raise ValueError(f'{fp.pid}: layers disagree on the number of windows')
            for w in wins:
                t = _as_cuda(w, dev)
                if t.dim() != 2:
                    raise ValueError('embeddings must be [rows, D]')
                D = t.shape[1] if D is None else D
                if t.shape[1] != D:
                    raise ValueError('all embeddings of a batch must share D')
                src[li].append(t)
        rows = [int(src[0][prot_src0[-1] + c].shape[0]) for c in range(len(wins0))]
        for li in range(1, n_layers):
            if [int(src[li][prot_src0[-1] + c].shape[0]) for c in range(len(wins0))] != rows:
                raise ValueError(f'{fp.pid}: layers disagree on the number of rows')
        src_rows += rows
        prot_len.append(rows[0] if len(rows) == 1 else (len(rows) - 1) * (maxlen - overlap) + rows[-1])

    dom_prot, dom_seg_off, seg_beg, seg_end, entries = [], [0], [], [], []
    for pi, fp in enumerate(fps):
        mine = []
        for dom in fp.domains:
            segs, kept = parse_domain(dom, prot_len[pi])
            if sum(e - b for b, e in segs) == 0:      # reference: empty embedding -> domain skipped
                continue
            mine.append((kept, len(dom_prot)))
            dom_prot.append(pi)
            for b, e in segs:
                seg_beg.append(b)
                seg_end.append(e)
            dom_seg_off.append(len(seg_beg))
        entries.append(mine)

    # ---- one launch per distinct (n, m) (the reference call site uses one: [3, 80, 3, 80]) ----
    blocks = []           # per layer: (tensor, column offset)
    groups: dict = {}
    for li in range(n_layers):
        groups.setdefault((int(qdim[2 * li]), int(qdim[2 * li + 1])), []).append(li)
    outs = {}
    if dom_prot:
        for (n, m), lids in groups.items():
            plan = make_plan(len(lids), D, n, m, src_rows, prot_src0, prot_nsrc, dom_prot, dom_seg_off,
                             seg_beg, seg_end, maxlen, overlap)
            out = torch.empty((len(dom_prot), len(lids) * n * m), dtype=torch.int8, device=dev)
            execute_plan(plan, [src[li] for li in lids], out)
            outs[(n, m)] = (out, lids)
        host = {key: (o.cpu().numpy(), lids) for key, (o, lids) in outs.items()}
        for li in range(n_layers):
            key = (int(qdim[2 * li]), int(qdim[2 * li + 1]))
            arr, lids = host[key]
            nm = key[0] * key[1]
            blocks.append(arr[:, lids.index(li) * nm:(lids.index(li) + 1) * nm])

    # ---- quants dicts, same update order as src/fingerprint.py:184-201 ----
    for pi, fp in enumerate(fps):
        for li in range(n_layers):
            for kept, row in entries[pi]:
                fp.quants.setdefault(kept, []).extend(blocks[li][row].tolist())
        for key, value in fp.quants.items():
            fp.quants[key] = np.array(value)
        fp.domains = list(fp.quants.keys())
    return fps


@dataclass
class Fingerprint:
    """Mirror of the reference dataclass (src/fingerprint.py:17-35)."""
    pid: str = field(default_factory=str)
    seq: str = field(default_factory=str)
    embed: dict = field(default_factory=dict)
    contacts: np.ndarray = field(default_factory=list)
    domains: list = field(default_factory=list)
    quants: dict = field(default_factory=dict)

    def __post_init__(self):
        self.contacts = np.array(self.contacts)

    # -- domain prediction stays as in the reference (RecCut is outside the accelerated path) --
    def writece(self, outfile: str, t: float):
        """Contact file for RecCut: top t*L contacts at sequence separation >= 5 (src/fingerprint.py:45-80)."""
        slen = len(self.seq)
        cta = self.contacts.reshape(slen, slen)
        iu, ju = np.triu_indices(slen, k=5)
        order = np.argsort(-cta[iu, ju], kind='stable')
        tot = min(int(t * slen), len(order))
        items = [f'{iu[o]} {ju[o]} {cta[iu[o]][ju[o]]:.6f}' for o in order[:tot]]
        sout = ('CON   ' + ','.join(items)) if items else ''
        with open(outfile, 'w', encoding='utf8') as out_f:
            out_f.write(f'INF   {self.pid} {slen}\n')
            out_f.write(f'SEQ   {self.seq}\n')
            out_f.write(f"SS    {'C' * slen}\n")
            out_f.write(sout + '\n')

    def reccut(self, threshold: float):
        """Runs the reference's RecCut binary (src/fingerprint.py:83-107).  The binary is looked up in
        $DCTD_RECCUT, then next to this file (``g++ -o RecCut RecCut.cpp`` as in the reference README)."""
        filename = f'{self.pid[:50]}.ce'
        self.writece(filename, threshold)
        rec_path = os.environ.get('DCTD_RECCUT') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'RecCut')
        try:
            result = sp.run([rec_path, '--input', filename, '--name', f'{self.pid}'],
                            stdout=sp.PIPE, text=True, check=True)
        finally:
            if os.path.exists(filename):
                os.remove(filename)
        domains = result.stdout.strip().split()[2].split(';')[:-1]
        self.domains.extend(domains)
        if len(domains) > 1:
            self.domains.append(f'1-{len(self.seq)}')

    def scale(self, vec: np.ndarray) -> np.ndarray:
        """(vec - min) / (max - min) over the whole array, float64 (src/fingerprint.py:110-123); on the GPU."""
        dev = _device()
        x = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(dev)
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            rc = _lib.lib().dctd_scale_f64(x.data_ptr(), x.numel(), out.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, 'dctd_scale_f64')
        return out.cpu().numpy().reshape(np.shape(vec))

    def idct_quant(self, vec: np.ndarray, num: int) -> np.ndarray:
        """iDCTquant of src/fingerprint.py:126-142 for an arbitrary [rows, cols] matrix: DCT-II (ortho) along
        the rows, first ``num`` coefficients, length-``num`` inverse, per-column min-max; returns
        [num, cols] float64.  Runs on the GPU (dctd_idct_quant_f64); ``quantize`` does not call it - the
        batched kernel fuses both passes."""
        dev = _device()
        x = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(dev)
        rows, cols = x.shape
        n_out = min(int(num), rows)
        out = torch.empty((n_out, cols), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().dctd_idct_quant_f64(x.data_ptr(), rows, cols, int(num), out.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, 'dctd_idct_quant_f64')
        return out.cpu().numpy()

    def get_doms(self, embed, dom: str):
        """Rows of one domain, float64, in listed order, and the kept string (src/fingerprint.py:145-171)."""
        segs, kept = parse_domain(dom, embed.shape[0])
        rows = [np.asarray(embed[b:e], dtype=np.float64) for b, e in segs]
        out = np.concatenate(rows, axis=0) if rows else np.empty((0, embed.shape[1]))
        return out, kept

    def quantize(self, qdim: list):
        """quant2D on the GPU (src/fingerprint.py:174-201)."""
        quantize_batch([self], qdim)


# the names make_db.py imports / calls (src/make_db.py:19-51), re-expressed over the batched call
def queue_cpu(fp: Fingerprint) -> Fingerprint:
    fp.reccut(2.6)
    fp.quantize([3, 80, 3, 80])
    return fp
