"""zelll-b200: a B200-native cell-list engine behind zelll's API (hot path only).

`CellGrid` mirrors `zelll.CellGrid` (python/src/lib.rs) and the Rust `CellGrid` entry points
(src/cellgrid.rs); all computation runs in libzelll_b200.so (hand-written sm_100a CUDA behind the
C ABI of include/zelll_b200.h).  There is no CPU fallback.
"""
from ._ffi import CMP_LE, CMP_LT, CMP_NONE, ZelllB200Error  # noqa: F401
from .cellgrid import CellGrid, CellGridIter, CellQueryIter, GridCell, GridInfo  # noqa: F401

__all__ = ["CellGrid", "CellGridIter", "CellQueryIter", "GridCell", "GridInfo", "ZelllB200Error", "CMP_NONE", "CMP_LT", "CMP_LE"]
