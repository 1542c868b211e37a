"""hubbardtn_b200 — B200-native (sm_100a) hot path behind HubbardTN's MPSKit calls.

Only what the path needs: `csrc/` (CUDA kernels + C ABI, built into libhtn.so), the ctypes
binding (`_lib`), the host-side object layer (`device`) and input builders (`sectors`,
`synthetic`).  Importing the package loads libhtn.so and fails loudly if it is missing.
"""
from . import _lib  # noqa: F401  (loads libhtn.so; raises ImportError when absent)
from . import sectors, synthetic, device  # noqa: F401

__all__ = ["sectors", "synthetic", "device"]
