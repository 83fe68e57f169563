"""TEST INFRASTRUCTURE ONLY — ctypes loader for the oracle's C restatement (oracle/alias_vose.c)."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmap_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "alias_vose.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmap_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def available() -> bool:
    global _lib
    if _lib is not None:
        return True
    if not os.path.exists(_SO):
        try:
            build()
        except Exception:
            return False
    try:
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_alias_build.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        _lib.oracle_alias_build.restype = ctypes.c_int
    except OSError:
        _lib = None
        return False
    return True


def alias_build(probs: torch.Tensor):
    assert available()
    p = probs.detach().cpu().float().contiguous()
    K = p.numel()
    out_prob = torch.empty(K, dtype=torch.float32)
    out_alias = torch.empty(K, dtype=torch.int64)
    rc = _lib.oracle_alias_build(p.data_ptr(), K, out_prob.data_ptr(), out_alias.data_ptr())
    if rc != 0:
        raise MemoryError("oracle_alias_build failed")
    return out_prob, out_alias
