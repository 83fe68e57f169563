"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/map_b200.h declares; host-only entry points (alias builder) are bit-exact against the reference goldens.
No device compute is issued here."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from map_code_b200 import _lib
    return _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "map_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(map_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_and_library_exports_every_symbol(lib):
    names = header_functions()
    assert len(names) >= 30
    cdll = lib.load()
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/map_b200.h but not exported by libmap_b200.so"
    assert sorted(lib.PROTOTYPES.keys()) == names, "ctypes prototype table out of sync with the header"
    assert cdll.map_abi_version() == 1


def test_struct_layouts_match_header(lib):
    import ctypes as C
    assert C.sizeof(lib.GemmArgs) == 24 + 8 * 13 + 8 * 4 + 8 + 8  # + aux2/acc_out (+ld), acc_accumulate/reserved, colsum_out
    assert C.sizeof(lib.AdamwTensor) == 80  # + planes pointer and plane stride (bf16 planes of the parameter)
    assert C.sizeof(lib.GemmSplitArgs) == C.sizeof(lib.GemmArgs) + 3 * 32
    assert C.sizeof(lib.HeadBwdArgs) == 256 and lib.HeadBwdArgs.scalar_col.offset == 228   # MapHeadBwdArgs (by-field encoder fold)


def test_errors_are_reported_not_swallowed(lib):
    with pytest.raises(lib.MapB200Error) as e:
        lib.call("map_alias_build", None, 0, None, None)
    assert "map_alias_build" in str(e.value)


def test_no_cpu_fallback_for_device_ops():
    from map_code_b200 import ops, _lib
    with pytest.raises(_lib.MapB200Error):
        ops.emb_gather(torch.zeros(4, 4), torch.zeros(2, dtype=torch.int64))


def test_alias_build_native_bit_exact(golden):
    from map_code_b200 import ops
    for case in ("kat", "zipf3000"):
        g = golden("alias")[case]
        prob, alias = ops.alias_build(g["probs"])
        assert torch.equal(alias, g["alias"]) and torch.equal(prob, g["prob"])


def test_alias_build_large_matches_oracle_c():
    """V = 1 085 271 (Criteo shape): the reference's Python loop needs ~35 s; both native builders agree bit for bit."""
    from map_code_b200 import ops
    from oracle import map_oracle as O
    g = torch.Generator().manual_seed(0)
    fc = torch.floor(torch.pow(torch.tensor(1e6), torch.rand(1_085_271, generator=g)))
    probs, _, _ = O.nce_noise_distribution(fc)
    p1, a1 = ops.alias_build(probs)
    p2, a2 = O.alias_build(probs)
    assert torch.equal(p1, p2) and torch.equal(a1, a2)
    # the table encodes the distribution: sum_k [prob_k + sum_{j: alias_j = k} (1 - prob_j)] / V == probs_k
    recon = p1.double().clone()
    recon.index_add_(0, a1, 1.0 - p1.double())
    err = (recon / probs.numel() - probs.double()).abs()
    assert err.max() < 1e-7 and err.sum() < 1e-3  # float32 Vose round-off, inherited bit-for-bit from the reference's loop


def test_alias_build_ignores_default_device_context():
    """bench.py builds the model under `with torch.device(cuda)`; the HOST table builder must still allocate on the CPU
    (a device allocation there handed a device pointer to host code: SIGSEGV at 2 GPUs, round 1)."""
    import torch
    from map_code_b200 import ops
    probs = torch.arange(1, 17, dtype=torch.float32) / 136.0
    with torch.device("meta"):
        prob, alias = ops.alias_build(probs)
    assert prob.device.type == "cpu" and alias.device.type == "cpu"
    assert alias.tolist() == [11, 13, 14, 14, 15, 15, 15, 15, 0, 8, 9, 10, 11, 12, 13, 14]  # SURVEY.md §8c KAT


def test_build_hook_and_makefile_compile_every_cuda_source():
    """__graft_entry__.build() and csrc/Makefile must both compile every .cu under csrc/ (a hand-written list once missed two files and
    a fresh checkout then failed to load the library)."""
    import os
    import re
    import __graft_entry__ as g
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "map_code_b200", "csrc")
    on_disk = sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))
    assert sorted(g.SOURCES) == on_disk
    mk = open(os.path.join(csrc, "Makefile")).read()
    srcs = re.search(r"^SRCS\s*:=\s*(.*)$", mk, re.M).group(1).split()
    assert sorted(srcs) == on_disk
