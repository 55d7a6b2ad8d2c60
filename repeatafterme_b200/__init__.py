"""repeatafterme_b200 -- B200-native banded-DP extension path of RepeatAfterMe's RAMExtend.

Only the hot path named by BASELINE.json lives here: the CUDA kernels + extern "C" ABI
(``csrc/``, built into ``libramcu.so``), the host-side C boundary that replaces the reference's
``extend_alignment()`` (``csrc/ramx_boundary.c``), and a thin ctypes mirror used by the parity
tests and ``bench.py``. Everything else of RAMExtend (CLI, loader, report, outputs) stays the
reference's own host C.
"""
from .abi import RamCu, load_ramcu, CORE_DTYPE, Scoring, ExtendResult  # noqa: F401
