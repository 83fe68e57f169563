"""map_code_b200 — B200-native (sm_100a) implementation of MAP's pretraining / finetuning hot path.

Mirrors the reference's Python surface (CHIANGEL/MAP-CODE `code/`): `layers`, `models`, `nce`, `trainer`, `arguments`.
Every device operation is a hand-written CUDA kernel in `csrc/` reached through the C ABI of `include/map_b200.h`
(`libmap_b200.so`, bound with ctypes in `_lib.py`).  There is no CPU, Triton or eager fallback.
"""
__version__ = "0.1.0"
