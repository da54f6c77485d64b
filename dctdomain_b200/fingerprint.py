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


def _owner(device):
    """Key of the per-(device, stream, host thread) scratch caches below: kernels queued on one stream by one thread
    run in order, so reusing a buffer is safe there and nowhere else (a search on one stream and fingerprinting on
    another, or two host threads, get separate scratch memory)."""
    import threading
    return device, torch.cuda.current_stream(device).cuda_stream, threading.get_ident()


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('dctdomain_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(device)


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Scratch memory of the calling (device, stream, thread).  Growing it replaces the buffer: anything an earlier call
    left in it (plan tables of ``execute_plan``) is gone, which is why ``tables_resident`` needs a caller-owned workspace."""
    key = _owner(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def parse_domain(dom: str, n_rows: int):
    """Rows of a RecCut domain string, with get_doms' bound handling (src/fingerprint.py:160-171).

    Returns ([(begin0, end_exclusive), ...], kept_string).  A segment is dropped when its begin
    lies beyond the protein, an end beyond the protein is clipped, and - as in the reference,
    which removes (by value) from the list it iterates over - the segment after a dropped one is
    passed over but stays in the returned string.
    """
    parts = dom.split(',')
    segs = []
    pos = 0
    while pos < len(parts):
        beg_s, end_s = parts[pos].split('-')
        beg, end = int(beg_s), int(end_s)
        if (beg or end) > n_rows:
            parts.remove(parts[pos])        # by value, like the reference: an equal segment passed over earlier goes first
            pos += 1
            continue
        lo = beg - 1 if beg >= 1 else n_rows + beg - 1
        hi = min(end, n_rows)
        segs.append((lo, max(lo, hi)))
        pos += 1
    return segs, ','.join(parts)


class _Plan:
    """Owns a dctd_fp_plan handle."""

    def __init__(self, geo: _lib.FpGeometry, keep, flags: int = 0):
        self._keep = keep
        self.handle = C.c_void_p()
        _lib.check(_lib.lib().dctd_fp_plan_create_ex(C.byref(geo), int(flags), C.byref(self.handle)), 'dctd_fp_plan_create')
        L = _lib.lib()
        self.workspace_bytes = int(L.dctd_fp_workspace_bytes(self.handle))
        self.algorithmic_bytes = int(L.dctd_fp_algorithmic_bytes(self.handle))
        self.n_items = int(L.dctd_fp_num_items(self.handle))

    def __del__(self):
        try:
            if getattr(self, 'handle', None) and self.handle.value:
                _lib.lib().dctd_fp_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:       # interpreter shutdown: the module globals may be gone already
            pass


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def make_plan(n_layers, D, n, m, src_rows, prot_src0, prot_nsrc, dom_prot, dom_seg_off, seg_beg, seg_end,
              maxlen=DEFAULT_MAXLEN, overlap=OVERLAP, flags: int = 0) -> _Plan:
    """Host-side work decomposition for one batch (dctd_fp_plan_create_ex; ``flags``: _lib.FP_PLAN_*)."""
    arrs = [_i32(x) for x in (src_rows, prot_src0, prot_nsrc, dom_prot, dom_seg_off, seg_beg, seg_end)]
    geo = _lib.FpGeometry()
    geo.n_layers, geo.D, geo.n, geo.m = int(n_layers), int(D), int(n), int(m)
    geo.maxlen, geo.overlap = int(maxlen), int(overlap)
    geo.n_src, geo.src_rows = len(arrs[0]), arrs[0].ctypes.data
    geo.n_prot, geo.prot_src0, geo.prot_nsrc = len(arrs[1]), arrs[1].ctypes.data, arrs[2].ctypes.data
    geo.n_dom, geo.dom_prot, geo.dom_seg_off = len(arrs[3]), arrs[3].ctypes.data, arrs[4].ctypes.data
    geo.seg_beg, geo.seg_end = arrs[5].ctypes.data, arrs[6].ctypes.data
    return _Plan(geo, arrs, flags)


def execute_plan(plan: _Plan, src_tensors, out: torch.Tensor, tables_resident=False, workspace=None):
    """Launch the fingerprint kernel.  ``src_tensors[layer][s]`` are CUDA float32 [rows, D] tensors.

    Stream-ordered on the current stream.  Without ``workspace`` the scratch memory of the calling (device, stream,
    thread) is used, which any other call on that stream may overwrite or replace: ``tables_resident=True`` (the plan
    tables were uploaded by an earlier call with this plan) therefore needs a caller-owned ``workspace``."""
    if tables_resident and workspace is None:
        raise ValueError('tables_resident=True needs the caller-owned workspace the plan tables were uploaded to')
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


def _host_f32(x):
    """numpy / CPU torch [rows, D] -> (keep-alive object, address, rows, D, pinned) of contiguous float32 host memory."""
    if isinstance(x, torch.Tensor):
        t = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()
        return t, t.data_ptr(), int(t.shape[0]), int(t.shape[1]), t.is_pinned()
    a = np.ascontiguousarray(x, dtype=np.float32)
    return a, a.ctypes.data, int(a.shape[0]), int(a.shape[1]), False


_GATHER_PIECE = 256 << 10      # bytes per piece of the gather kernel's work list
_tables: dict = {}


def _pinned_table(device, n_entries: int) -> np.ndarray:
    """Pinned host array [n_entries, 3] of int64 (src, dst, nbytes descriptors the gather kernel reads over PCIe)."""
    key = _owner(device)
    t = _tables.get(key)
    if t is None or t.shape[0] < n_entries:
        t = torch.empty((max(n_entries, 4096) * 2, 3), dtype=torch.int64).pin_memory()
        _tables[key] = t
    return t


_RING_SLOT = 4 << 20           # bytes per slot of the pinned ring pageable arrays are staged through
_RING_SLOTS = 16
_rings: dict = {}


def _pinned_ring(device) -> torch.Tensor:
    key = _owner(device)
    r = _rings.get(key)
    if r is None:
        r = _rings[key] = torch.empty(_RING_SLOT * _RING_SLOTS, dtype=torch.uint8).pin_memory()
    return r


def _stage_threads() -> int:
    """Host threads that copy pageable arrays into the pinned ring: DCTD_STAGE_THREADS, else half the CPUs this process may
    run on, shared between the ranks of the box (LOCAL_WORLD_SIZE), at most 8."""
    env = os.environ.get('DCTD_STAGE_THREADS')
    if env:
        return max(1, min(64, int(env)))
    try:
        cpus = len(os.sched_getaffinity(0))
    except AttributeError:
        cpus = os.cpu_count() or 2
    ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1') or 1))
    return max(1, min(8, cpus // (2 * ranks)))


_stage: dict = {}
_aux: dict = {}


def _aux_stream(device) -> torch.cuda.Stream:
    key = _owner(device)
    st = _aux.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _aux[key] = st
    return st


def _stage_buffer(device, nbytes):
    key = _owner(device)
    buf = _stage.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + (1 << 20), dtype=torch.uint8, device=device)
        _stage[key] = buf
    return buf


def parse_domains_batch(dom_lists, prot_len):
    """All RecCut strings of a batch -> (dom_prot, dom_seg_off, seg_beg, seg_end, names): int32 arrays for ``make_plan``
    and, per output domain, the (kept) string.  One C call (dctd_parse_domains) for the regular strings; a batch that
    holds a string needing get_doms' special paths (src/fingerprint.py:163-165) goes through ``parse_domain``, the
    line-by-line mirror of the reference, so the result is the same either way."""
    counts = [len(d) for d in dom_lists]
    strs = [s for d in dom_lists for s in d]
    n_str = len(strs)
    plen = _i32(prot_len)
    if n_str:
        text = ('\n'.join(strs) + '\n').encode()
        if text.count(b'\n') == n_str:
            str_prot = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
            max_segs = text.count(b',') + n_str
            dom_str = np.empty(n_str, dtype=np.int32)
            seg_off = np.empty(n_str + 1, dtype=np.int32)
            seg_beg = np.empty(max_segs, dtype=np.int32)
            seg_end = np.empty(max_segs, dtype=np.int32)
            n_dom, n_irr = C.c_int32(), C.c_int32()
            _lib.check(_lib.lib().dctd_parse_domains(text, len(text), n_str, str_prot.ctypes.data, plen.ctypes.data, len(plen),
                                                     dom_str.ctypes.data, seg_off.ctypes.data, seg_beg.ctypes.data,
                                                     seg_end.ctypes.data, max_segs, C.byref(n_dom), C.byref(n_irr)),
                       'dctd_parse_domains')
            if n_irr.value == 0:
                nd = n_dom.value
                ns = int(seg_off[nd])
                ds = dom_str[:nd]
                names = strs if nd == n_str else [strs[i] for i in ds.tolist()]
                return str_prot[ds], seg_off[:nd + 1], seg_beg[:ns], seg_end[:ns], names
    dom_prot, dom_seg_off, seg_beg, seg_end, names = [], [0], [], [], []
    for pi, doms in enumerate(dom_lists):
        for dom in doms:
            segs, kept = parse_domain(dom, int(plen[pi]))
            if sum(e - b for b, e in segs) == 0:      # reference: empty embedding -> domain skipped
                continue
            names.append(kept)
            dom_prot.append(pi)
            for b, e in segs:
                seg_beg.append(b)
                seg_end.append(e)
            dom_seg_off.append(len(seg_beg))
    return _i32(dom_prot), _i32(dom_seg_off), _i32(seg_beg), _i32(seg_end), names


def _walk_cuda(fps, dev, n_layers, maxlen, overlap):
    """Sources of a batch whose embeddings are all ready-to-use CUDA tensors (float32, contiguous, on ``dev``): device
    addresses per layer and the source geometry, with the fewest attribute reads per tensor.  None if anything else
    turns up (host arrays, other dtypes ...): the general walk then handles - and converts - it."""
    T = torch.Tensor
    f32 = torch.float32
    idx = dev.index
    ptr = [[] for _ in range(n_layers)]
    src_rows, prot_src0, prot_nsrc, prot_len = [], [], [], []
    D = None
    stride = maxlen - overlap
    for fp in fps:
        vals = list(fp.embed.values())
        if len(vals) != n_layers:
            raise ValueError('all proteins of a batch must carry the same layers')
        v0 = vals[0]
        if type(v0) is T:
            shape0 = None
            for li in range(n_layers):
                w = vals[li]
                if type(w) is not T or not w.is_cuda or w.dtype is not f32 or w.get_device() != idx or not w.is_contiguous():
                    return None
                sh = w.shape
                if shape0 is None:
                    if len(sh) != 2:
                        raise ValueError('embeddings must be [rows, D]')
                    shape0 = sh
                    if D is None:
                        D = sh[1]
                    elif sh[1] != D:
                        raise ValueError('all embeddings of a batch must share D')
                elif sh != shape0:
                    raise ValueError(f'{fp.pid}: layers disagree on the number of rows')
                ptr[li].append(w.data_ptr())
            prot_src0.append(len(src_rows))
            prot_nsrc.append(1)
            src_rows.append(shape0[0])
            prot_len.append(shape0[0])
        elif isinstance(v0, (list, tuple)):
            nwin = len(v0)
            rows0 = None
            for li in range(n_layers):
                wins = vals[li]
                if not isinstance(wins, (list, tuple)):
                    return None
                if len(wins) != nwin:
                    raise ValueError(f'{fp.pid}: layers disagree on the number of windows')
                rows = []
                for w in wins:
                    if type(w) is not T or not w.is_cuda or w.dtype is not f32 or w.get_device() != idx or not w.is_contiguous():
                        return None
                    sh = w.shape
                    if len(sh) != 2:
                        raise ValueError('embeddings must be [rows, D]')
                    if D is None:
                        D = sh[1]
                    elif sh[1] != D:
                        raise ValueError('all embeddings of a batch must share D')
                    rows.append(sh[0])
                    ptr[li].append(w.data_ptr())
                if rows0 is None:
                    rows0 = rows
                elif rows != rows0:
                    raise ValueError(f'{fp.pid}: layers disagree on the number of rows')
            prot_src0.append(len(src_rows))
            prot_nsrc.append(nwin)
            src_rows += rows0
            prot_len.append(rows0[0] if nwin == 1 else (nwin - 1) * stride + rows0[-1])
        else:
            return None
    return ptr, src_rows, prot_src0, prot_nsrc, prot_len, D


@dataclass
class DeviceFingerprints:
    """Result of ``quantize_device``: ``fingerprints`` int8 CUDA tensor [n_dom, n_layers*n*m] (layer-major, the layout of
    src/fingerprint.py:194-200), ``dom_prot`` protein of each row, ``names`` its (kept) RecCut string."""
    fingerprints: torch.Tensor
    dom_prot: np.ndarray
    names: list


def quantize_device(layers, row_start, row_count, domains, qdim=(3, 80, 3, 80), out=None, plan_flags: int = 0):
    """``quantize`` for embeddings that stay on the device as whole-batch tensors - the form an ESM-2 forward pass leaves
    them in - with no per-protein Python work:

      layers      one CUDA float32 tensor [rows_total, D] per embedding layer (e.g. the padded model output
                  [B, T, D] viewed as [B*T, D]); all layers share the geometry
      row_start,  protein p is rows [row_start[p], row_start[p] + row_count[p]) of every layer tensor
      row_count   (for a padded batch: p*T + 1 and the sequence length: the BOS / EOS rows are simply not addressed)
      domains     per protein the list of RecCut strings (what ``Fingerprint.domains`` holds), parsed with get_doms'
                  rules in one C call

    One kernel launch per distinct (n, m) of ``qdim``; stream-ordered, nothing is copied to the host.  Returns
    ``DeviceFingerprints``; rows are in the order of ``domains`` (strings without rows are skipped, as the reference
    does).  Proteins delivered as several maxlen windows go through ``quantize_batch``."""
    dev = _device(layers[0].device)
    n_layers = len(layers)
    qdim = list(qdim)
    if len(qdim) < 2 * n_layers:
        raise IndexError('qdim needs an (n, m) pair per embedding layer')
    D = int(layers[0].shape[1])
    for t in layers:
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous() and t.shape == layers[0].shape
                and t.device == dev):
            raise ValueError('layers must be contiguous float32 CUDA tensors [rows_total, D] of one shape on one device')
    row_start = np.ascontiguousarray(row_start, dtype=np.int64)
    row_count = _i32(row_count)
    n_prot = len(row_count)
    if len(row_start) != n_prot or len(domains) != n_prot:
        raise ValueError('row_start, row_count and domains need one entry per protein')
    if n_prot and (row_start.min() < 0 or (row_start + row_count).max() > layers[0].shape[0] or row_count.min() < 0):
        raise ValueError('protein rows outside the layer tensors')
    dom_prot, seg_off, seg_beg, seg_end, names = parse_domains_batch(domains, row_count)
    nd = len(dom_prot)
    groups: dict = {}
    for li in range(n_layers):
        groups.setdefault((int(qdim[2 * li]), int(qdim[2 * li + 1])), []).append(li)
    width = sum(int(qdim[2 * li]) * int(qdim[2 * li + 1]) for li in range(n_layers))
    if out is None:
        out = torch.empty((nd, width), dtype=torch.int8, device=dev)
    elif out.shape != (nd, width) or out.dtype != torch.int8 or out.device != dev or out.stride(1) != 1:
        raise ValueError(f'out must be an int8 CUDA tensor [{nd}, {width}]')
    if nd == 0:
        return DeviceFingerprints(out, dom_prot, names)
    L = _lib.lib()
    stream = torch.cuda.current_stream(dev).cuda_stream
    ar = np.arange(n_prot, dtype=np.int32)
    ones = np.ones(n_prot, dtype=np.int32)
    col = np.concatenate([[0], np.cumsum([int(qdim[2 * li]) * int(qdim[2 * li + 1]) for li in range(n_layers)])])
    for (n, m), lids in groups.items():
        if lids != list(range(lids[0], lids[0] + len(lids))):
            raise ValueError('layers that share an (n, m) pair must be adjacent in qdim')     # their columns are one block
        plan = make_plan(len(lids), D, n, m, row_count, ar, ones, dom_prot, seg_off, seg_beg, seg_end, flags=plan_flags)
        ptrs = np.concatenate([np.uint64(layers[li].data_ptr()) + (row_start * (D * 4)).astype(np.uint64) for li in lids])
        ws = _workspace(dev, plan.workspace_bytes)
        view = out[:, int(col[lids[0]]):]
        with torch.cuda.device(dev):
            _lib.check(L.dctd_fp_execute(plan.handle, ptrs.ctypes.data, D, view.data_ptr(), out.stride(0),
                                         ws.data_ptr(), ws.numel(), 0, stream), 'dctd_fp_execute')
        _retire(dev, plan)
    return DeviceFingerprints(out, dom_prot, names)


_retired: dict = {}


def _retire(device, plan):
    """Keeps a plan alive until the launch that uploads its tables has certainly consumed them: plans are destroyed two
    calls later on the same (device, stream, thread), and dctd_fp_plan_destroy itself waits for the upload event."""
    key = _owner(device)
    q = _retired.setdefault(key, [])
    q.append(plan)
    if len(q) > 4:
        del q[0]


def quantize_batch(fps, qdim=(3, 80, 3, 80), device=None, maxlen=DEFAULT_MAXLEN, overlap=OVERLAP, plan_flags: int = 0,
                   quants_dtype=np.int64, staging: str = 'auto', _timing=None, _copy_stream=None, _turn=None):
    """``quantize`` for a list of Fingerprint-like objects in one kernel launch per (n, m) group.

    Each object needs ``embed`` ({layer: [L, D] array | list of window arrays}), ``domains`` (list of
    RecCut strings) and ``quants`` (dict); they are updated exactly as the reference ``quantize`` does
    (src/fingerprint.py:184-201).  Host embeddings (numpy, CPU torch, pinned or not) are staged into one
    device buffer with a single C call (one cudaMemcpyAsync per array); CUDA tensors are read in place.
    ``plan_flags``: per-call options of the work decomposition (``_lib.FP_PLAN_NO_FUSION`` ...).  ``quants_dtype``: dtype
    of the arrays put into ``quants`` - int64 is what the reference produces (``np.array`` of Python ints,
    src/fingerprint.py:200); ``np.int8`` skips the widening (same values, an eighth of the bytes).  ``staging``: how
    host arrays reach the device - 'auto' / 'gather' (pinned arrays: one gather kernel pulling all arrays over PCIe;
    pageable arrays: host threads copy them into a pinned ring, one copy-engine transfer per 4 MB slot) or 'dma' (one
    cudaMemcpyAsync per array, whatever the memory).  Returns the list.
    """
    fps = list(fps)
    if not fps:
        return fps
    dev = _device(device)
    n_layers = len(fps[0].embed)
    qdim = list(qdim)
    if len(qdim) < 2 * n_layers:
        raise IndexError('qdim needs an (n, m) pair per embedding layer')  # reference: list index error
    L = _lib.lib()

    fast = _walk_cuda(fps, dev, n_layers, maxlen, overlap)
    keep, h_addr = [], []
    main_stream = torch.cuda.current_stream(dev)
    stream = main_stream.cuda_stream
    if fast is not None:
        # every embedding is a ready-to-use CUDA tensor (the fps keep them alive): nothing to stage
        ptr, src_rows, prot_src0, prot_nsrc, prot_len, D = fast
    else:
        # ---- sources: device pointer per (layer, source); host arrays are staged ----
        ptr = [[] for _ in range(n_layers)]        # device addresses (staged ones are offsets until the copy)
        staged = [[] for _ in range(n_layers)]     # True where ptr holds an offset into the staging buffer
        keep, h_addr, h_bytes, h_off, h_pin = [], [], [], [], []
        src_rows, prot_src0, prot_nsrc, prot_len = [], [], [], []
        D = None
        stage_bytes = 0
        main_stream = torch.cuda.current_stream(dev)
        stream = main_stream.cuda_stream
        # host arrays arrive as many separate allocations; they are staged on a side stream (batched submission,
        # dctd_h2d_rows) while this thread keeps walking the batch, the compute stream joins before the kernel
        # (quantize_stream hands in one copy stream for all batches in flight and a turn: batches queue their copies one
        # after the other, so the link finishes batch i before it starts on batch i + 1 instead of sharing itself out)
        aux_stream = _copy_stream if _copy_stream is not None else _aux_stream(dev)
        aux_stream.wait_stream(main_stream)
        have = _stage.get(_owner(dev))    # staging buffer of an earlier call (kept alive until this call returns)
        state = {'done': 0, 'bases': [], 'tab': 0}  # bases: [(first h index, device address that h_off is relative to)]

        def flush(final=False):
            """Issues the H2D copies collected so far.  While the staging buffer of an earlier call is large enough
            the copies start while the batch is still being walked (DMA overlaps the Python loop); whatever does not
            fit goes to a buffer allocated once the total is known."""
            lo = state['done']
            if lo == len(h_addr):
                return
            if not final and staging != 'dma' and not all(h_pin[lo:]):
                return                  # pageable arrays go in one staged call at the end (host threads per call)
            if _turn is not None and not _turn[0].is_now(_turn[1]):
                if not final:
                    return              # not this batch's turn yet: keep walking, the copies are issued later
                _turn[0].wait(_turn[1])
            if have is not None and stage_bytes <= have.numel():
                base = have.data_ptr()
                if not state['bases']:
                    state['bases'].append((0, base))
            elif final:
                first = h_off[lo]
                base = _stage_buffer(dev, stage_bytes).data_ptr() - first     # remaining copies, rebased
                state['bases'].append((lo, base))
            else:
                return
            if _timing is not None and 'copies_begin_event' not in _timing:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record(aux_stream)
                _timing['copies_begin_event'] = ev0
            a_src = np.array(h_addr[lo:], dtype=np.uint64)
            a_len = np.array(h_bytes[lo:], dtype=np.int64)
            a_off = np.array(h_off[lo:], dtype=np.int64)
            with torch.cuda.device(dev):
                if staging != 'dma' and all(h_pin[lo:]) and not (a_len % 16).any() and not (a_src % 16).any():     # 16-byte loads: sizes AND addresses
                    # pinned sources: one gather kernel over <= 256 KB pieces (no per-array DMA gaps)
                    npc = (a_len + _GATHER_PIECE - 1) // _GATHER_PIECE
                    total = int(npc.sum())
                    owner = np.repeat(np.arange(len(a_len)), npc)
                    first = np.cumsum(npc) - npc
                    within = (np.arange(total) - first[owner]) * _GATHER_PIECE
                    t0 = state['tab']
                    table = _pinned_table(dev, t0 + total)
                    view = table.numpy()[t0:t0 + total]
                    view[:, 0] = a_src[owner].astype(np.int64) + within
                    view[:, 1] = base + a_off[owner] + within
                    view[:, 2] = np.minimum(a_len[owner] - within, _GATHER_PIECE)
                    state['tab'] = t0 + total
                    keep.append(table)
                    _lib.check(L.dctd_h2d_gather(table.data_ptr() + t0 * 24, total, aux_stream.cuda_stream), 'dctd_h2d_gather')
                elif staging != 'dma':
                    # pageable arrays (numpy: what the reference's .cpu().numpy() leaves): host threads copy them into a
                    # pinned ring, one copy-engine transfer per slot (cudaMemcpyAsync from pageable memory: ~10 GB/s)
                    ring = _pinned_ring(dev)
                    _lib.check(L.dctd_h2d_rows_staged(a_src.ctypes.data, a_len.ctypes.data, len(a_src), base, a_off.ctypes.data,
                                                      ring.data_ptr(), _RING_SLOT, ring.numel() // _RING_SLOT, _stage_threads(),
                                                      aux_stream.cuda_stream), 'dctd_h2d_rows_staged')
                else:
                    _lib.check(L.dctd_h2d_rows(a_src.ctypes.data, a_len.ctypes.data, len(a_src), base,
                                               a_off.ctypes.data, aux_stream.cuda_stream), 'dctd_h2d_rows')
            state['done'] = len(h_addr)

        for fi, fp in enumerate(fps):
            if fi % 32 == 31 or fi in (2, 8):       # start the DMA early, then keep it fed
                flush()
            if len(fp.embed) != n_layers:
                raise ValueError('all proteins of a batch must carry the same layers')
            layers = list(fp.embed.values())
            nwin = len(layers[0]) if isinstance(layers[0], (list, tuple)) else 1
            prot_src0.append(len(src_rows))
            prot_nsrc.append(nwin)
            rows0 = None
            for li, lay in enumerate(layers):
                wins = lay if isinstance(lay, (list, tuple)) else [lay]
                if len(wins) != nwin:
                    raise ValueError(f'{fp.pid}: layers disagree on the number of windows')
                rows = []
                for w in wins:
                    if isinstance(w, torch.Tensor) and w.is_cuda:
                        t = w if (w.dtype == torch.float32 and w.is_contiguous() and w.device == dev) \
                            else w.to(dev, torch.float32).contiguous()
                        if t.dim() != 2:
                            raise ValueError('embeddings must be [rows, D]')
                        keep.append(t)
                        r, dd = int(t.shape[0]), int(t.shape[1])
                        ptr[li].append(t.data_ptr())
                        staged[li].append(False)
                    else:
                        if np.ndim(w) != 2:
                            raise ValueError('embeddings must be [rows, D]')
                        obj, addr, r, dd, pinned = _host_f32(w)
                        keep.append(obj)
                        h_pin.append(pinned)
                        h_addr.append(addr)
                        h_bytes.append(r * dd * 4)
                        h_off.append(stage_bytes)
                        ptr[li].append(len(h_addr) - 1)       # index into h_*; resolved after the copies are issued
                        staged[li].append(True)
                        stage_bytes += (r * dd * 4 + 255) // 256 * 256
                    D = dd if D is None else D
                    if dd != D:
                        raise ValueError('all embeddings of a batch must share D')
                    rows.append(r)
                if rows0 is None:
                    rows0 = rows
                elif rows != rows0:
                    raise ValueError(f'{fp.pid}: layers disagree on the number of rows')
            src_rows += rows0
            prot_len.append(rows0[0] if nwin == 1 else (nwin - 1) * (maxlen - overlap) + rows0[-1])

        flush(final=True)
        main_stream.wait_stream(aux_stream)
        if _turn is not None:
            _turn[0].wait(_turn[1])
            _turn[0].done(_turn[1])     # every copy of this batch is queued: the next batch may queue its own
        if h_addr:
            def resolve(hi):
                base = state['bases'][-1][1] if hi >= state['bases'][-1][0] else state['bases'][0][1]
                return base + h_off[hi]
            for li in range(n_layers):
                ptr[li] = [resolve(v) if st else v for v, st in zip(ptr[li], staged[li])]

    if _timing is not None:                 # debug hook (scripts/e2e_phases.py): host time stamps of the call's phases
        import time as _t
        _timing['walk_done'] = _t.perf_counter()
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(main_stream)
        _timing['copies_event'] = ev

    # ---- domains ----
    dom_prot, dom_seg_off, seg_beg, seg_end, names = parse_domains_batch([fp.domains for fp in fps], prot_len)
    if _timing is not None:
        _timing['parsed'] = _t.perf_counter()
    n_dom = len(dom_prot)

    # ---- one launch per distinct (n, m) (the reference call site uses one: [3, 80, 3, 80]) ----
    groups: dict = {}
    for li in range(n_layers):
        groups.setdefault((int(qdim[2 * li]), int(qdim[2 * li + 1])), []).append(li)
    blocks = [None] * n_layers       # per layer: int8 host array [n_dom, n*m]
    if n_dom:
        outs = []
        for (n, m), lids in groups.items():
            plan = make_plan(len(lids), D, n, m, src_rows, prot_src0, prot_nsrc, dom_prot, dom_seg_off,
                             seg_beg, seg_end, maxlen, overlap, plan_flags)
            out = torch.empty((n_dom, len(lids) * n * m), dtype=torch.int8, device=dev)
            ptrs = np.array([v for li in lids for v in ptr[li]], dtype=np.uint64)
            ws = _workspace(dev, plan.workspace_bytes)
            with torch.cuda.device(dev):
                _lib.check(L.dctd_fp_execute(plan.handle, ptrs.ctypes.data, D, out.data_ptr(), out.stride(0),
                                             ws.data_ptr(), ws.numel(), 0, stream), 'dctd_fp_execute')
            outs.append((out, n * m, lids, plan))
        if _timing is not None:
            _timing['launched'] = _t.perf_counter()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(main_stream)
            _timing['kernel_event'] = ev
            _timing['copies_event'].synchronize()
            _timing['copies_done'] = _t.perf_counter()
        # wait on an event first: Event.synchronize releases the GIL, so the other batches of quantize_stream keep
        # walking / assembling while this one is on the device (the blocking copy below then returns at once)
        done = torch.cuda.Event()
        done.record(main_stream)
        done.synchronize()
        for out, nm, lids, _plan in outs:
            arr = out.cpu().numpy()
            for pos, li in enumerate(lids):
                blocks[li] = arr[:, pos * nm:(pos + 1) * nm]

    if _timing is not None:
        _timing['results_on_host'] = _t.perf_counter()
    # ---- quants dicts, same update order as src/fingerprint.py:184-201 ----
    # the reference's values are int64 arrays of n*m entries per layer, layer after layer (fingerprint.py:194-200)
    rows, first = [], None
    if n_dom:
        wide = blocks[0] if n_layers == 1 else np.concatenate([blocks[li] for li in range(n_layers)], axis=1)
        if wide.dtype != quants_dtype:
            wide = wide.astype(quants_dtype)
        rows = list(wide)                                   # one view per domain
        first = np.searchsorted(dom_prot, np.arange(len(fps) + 1)).tolist()     # dom_prot is ascending
    for pi, fp in enumerate(fps):
        a, b = (first[pi], first[pi + 1]) if n_dom else (0, 0)
        mine = names[a:b]
        if not fp.quants and len(set(mine)) == len(mine):
            # the usual case: fresh object, every domain listed once (rows of one widened block)
            fp.quants.update(zip(mine, rows[a:b]))
        else:
            for li in range(n_layers):
                for kept, row in zip(mine, range(a, b)):
                    fp.quants.setdefault(kept, []).extend(blocks[li][row].tolist())
            for key, value in fp.quants.items():
                fp.quants[key] = np.array(value)
        fp.domains = list(fp.quants.keys())
    if not n_dom and h_addr:
        torch.cuda.current_stream(dev).synchronize()   # staged copies still read the host arrays
    del keep
    return fps


_stream_pool = None          # worker threads of quantize_stream (persistent: their scratch caches are reused)
_stream_local = None         # threading.local: one CUDA stream per worker thread and device
_copy_streams: dict = {}     # device -> the copy stream the batches of quantize_stream share


class _Turn:
    """Batches of one quantize_stream call queue their host-to-device copies in submission order."""

    def __init__(self):
        import threading
        self.cond = threading.Condition()
        self.now = 0

    def is_now(self, seq):
        return self.now == seq

    def wait(self, seq):
        with self.cond:
            while self.now < seq:
                self.cond.wait()

    def done(self, seq):
        with self.cond:
            if self.now == seq:
                self.now = seq + 1
                self.cond.notify_all()


def quantize_stream(batches, qdim=(3, 80, 3, 80), device=None, depth: int = 2, **kwargs):
    """``quantize_batch`` over an iterable of batches with ``depth`` batches in flight: the loop of ``make_db.main``
    (src/make_db.py:36-51: embed a batch, fingerprint it, store it) as a generator that yields the finished batches in
    order.  While the host arrays of one batch cross PCIe, the next batch is walked and planned and the previous one is
    turned into ``quants`` dicts, so the link never waits for the interpreter (a single ``quantize_batch`` call leaves it
    idle for ~2 ms of 30 per 512 proteins).  Every batch runs as one ordinary ``quantize_batch`` call on a worker thread
    with its own CUDA stream; the scratch caches are per (device, stream, thread) and the worker threads are kept, so the
    calls share nothing but the copy stream, on which the batches queue their copies one after the other (copies of two
    batches sharing the link would finish together and leave the same gap).  Results are the bytes ``quantize_batch``
    gives.  ``kwargs`` go to ``quantize_batch``; an exception of a batch is raised when that batch is due."""
    import collections
    import threading
    from concurrent.futures import ThreadPoolExecutor
    global _stream_pool, _stream_local
    if not 1 <= depth <= 4:
        raise ValueError('depth must be 1..4')
    dev = _device(device)
    if _stream_pool is None:
        _stream_local = threading.local()
        _stream_pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix='dctd-fp')
    copy_stream = _copy_streams.get(dev)
    if copy_stream is None:
        copy_stream = _copy_streams[dev] = torch.cuda.Stream(device=dev)
    turn = _Turn()

    def work(fps, seq):
        streams = getattr(_stream_local, 'by_device', None)
        if streams is None:
            streams = _stream_local.by_device = {}
        st = streams.get(dev)
        if st is None:
            # high priority: when the copies of the next batch (gather kernels on the default-priority copy stream)
            # fill the SMs, this batch's fingerprint kernel gets the next free slots instead of waiting for them to end
            st = streams[dev] = torch.cuda.Stream(device=dev, priority=-1)
        try:
            with torch.cuda.device(dev), torch.cuda.stream(st):
                return quantize_batch(fps, qdim, device=dev, _copy_stream=copy_stream, _turn=(turn, seq), **kwargs)
        finally:
            turn.wait(seq)          # a batch that raised, was empty or needed no copies still passes the turn on
            turn.done(seq)

    pending = collections.deque()
    for seq, fps in enumerate(batches):
        pending.append(_stream_pool.submit(work, fps, seq))
        if len(pending) >= depth:
            yield pending.popleft().result()
    while pending:
        yield pending.popleft().result()


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
