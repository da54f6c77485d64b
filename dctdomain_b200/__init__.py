"""dctdomain_b200 - B200 (sm_100a) implementation of DCTdomain's two data-parallel hot paths.

  fingerprint : drop-in for reference src/fingerprint.py (quantize on the GPU, batched)
  index       : drop-in for the faiss calls of src/database.py:241-243 / src/query_db.py:75-76,87
  query_db    : get_top_hits / search_db of src/query_db.py
  dct_sim     : src/dct-sim.py
  sharded     : database sharded over the GPUs of one box, NCCL merge of the per-rank top-k

Everything numeric runs in libdctd.so (hand-written CUDA, C ABI in include/dctd.h); there is no
CPU fallback.
"""
import sys

from . import _lib  # noqa: F401


def install_faiss_shim():
    """Registers dctdomain_b200.index as ``faiss`` so that the reference's ``import faiss``
    (src/database.py:8, src/query_db.py:8) binds to the GPU index without editing those files."""
    from . import index
    sys.modules['faiss'] = index
    return index


__all__ = ['_lib', 'install_faiss_shim']
