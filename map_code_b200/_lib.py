"""ctypes binding of ``libmap_b200.so`` (C ABI declared in ``include/map_b200.h``).

There is no fallback: if the library is missing or a call fails, this raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C map_code_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmap_b200.so")

MAP_OK, MAP_EINVAL, MAP_ECUDA, MAP_EWORKSPACE, MAP_EUNSUPPORTED = 0, -1, -2, -3, -4

# enums of the header
SCHED_CONST, SCHED_COSINE = 0, 1
SAMPLING_RANDINT, SAMPLING_NORMAL = 0, 1
RFD_MODES = {"Unigram": 0, "Uniform": 1, "Whole-Uniform": 2, "Whole-Unigram": 3}
NCE_LOSS = {"nce": 0, "sampled": 1}
EPI_NONE, EPI_BIAS, EPI_BIAS_RELU, EPI_CROSS, EPI_MUL_RELUMASK, EPI_ADD, EPI_ADD_MUL, EPI_CROSS_BWD, EPI_ADD3 = range(9)


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("trans_a", C.c_int32), ("trans_b", C.c_int32), ("epilogue", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("C", C.c_void_p), ("ldc", C.c_int64),
        ("bias", C.c_void_p),
        ("aux0", C.c_void_p), ("ld_aux0", C.c_int64),
        ("aux1", C.c_void_p), ("ld_aux1", C.c_int64),
        ("aux_out", C.c_void_p), ("ld_aux_out", C.c_int64),
        ("aux2", C.c_void_p), ("ld_aux2", C.c_int64),
        ("acc_out", C.c_void_p), ("ld_acc_out", C.c_int64),
        ("acc_accumulate", C.c_int32), ("reserved_", C.c_int32),
        ("colsum_out", C.c_void_p),
    ]


class AdamwTensor(C.Structure):
    _fields_ = [
        ("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("p_t", C.c_void_p),
        ("n", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32), ("weight_decay", C.c_float), ("n_planes", C.c_int32),
        ("planes", C.c_void_p), ("plane_stride", C.c_int64),
    ]


class GemmSplitArgs(C.Structure):
    _fields_ = [
        ("g", GemmArgs),
        ("a_planes", C.c_void_p), ("a_ld", C.c_int64), ("a_plane_stride", C.c_int64), ("a_nplanes", C.c_int32), ("terms", C.c_int32),
        ("b_planes", C.c_void_p), ("b_ld", C.c_int64), ("b_plane_stride", C.c_int64), ("b_nplanes", C.c_int32), ("reserved0_", C.c_int32),
        ("c_planes", C.c_void_p), ("c_ld", C.c_int64), ("c_plane_stride", C.c_int64), ("c_nplanes", C.c_int32), ("reserved1_", C.c_int32),
    ]


class HeadBwdArgs(C.Structure):
    _fields_ = [
        ("dxpos", C.c_void_p), ("ld_dx", C.c_int64),
        ("B", C.c_int64), ("L", C.c_int32), ("ncols", C.c_int32),
        ("cross_col0", C.c_int32), ("cross_w", C.c_int32),
        ("x0", C.c_void_p), ("ld_x0", C.c_int64),
        ("u", C.c_void_p), ("ld_u", C.c_int64),
        ("g_out", C.c_void_p), ("ld_g", C.c_int64),
        ("du_out", C.c_void_p), ("ld_du", C.c_int64),
        ("dx0_out", C.c_void_p), ("ld_dx0", C.c_int64),
        ("du_planes", C.c_void_p), ("du_pl_ld", C.c_int64), ("du_pl_stride", C.c_int64), ("du_nplanes", C.c_int32), ("reserved0_", C.c_int32),
        ("cross_bias_grad", C.c_void_p),
        ("relu_col0", C.c_int32), ("relu_w", C.c_int32),
        ("y", C.c_void_p), ("ld_y", C.c_int64),
        ("dz_out", C.c_void_p), ("ld_dz", C.c_int64),
        ("dz_planes", C.c_void_p), ("dz_pl_ld", C.c_int64), ("dz_pl_stride", C.c_int64), ("dz_nplanes", C.c_int32), ("scalar_col", C.c_int32),
        ("relu_bias_grad", C.c_void_p),
        ("scalar_out", C.c_void_p), ("ld_scalar", C.c_int64),
    ]


_p, _i, _l, _f, _u64, _sz, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t, C.c_double

# name -> (restype, argtypes).  Kept in the order of include/map_b200.h; tests/test_abi.py checks it against the header.
PROTOTYPES = {
    "map_abi_version": (_i, []),
    "map_last_error": (C.c_char_p, []),
    "map_sm_count": (_i, [C.POINTER(C.c_int)]),
    "map_timestamp_ns": (_i, [_p, _p]),
    "map_emb_gather_f32": (_i, [_p, _l, _i, _p, _l, _p, _p, _p]),
    "map_dedup_workspace_bytes": (_sz, [_l]),
    "map_dedup_ids": (_i, [_p, _l, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "map_segment_reduce_rows": (_i, [_p, _l, _i, _p, _i, _p, _p, _p, _l, _p, _p, _p]),
    "map_dedup_ids_ex": (_i, [_p, _l, _p, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "map_dedup_debug_offset": (_sz, [_l]),
    "map_segment_reduce_rows_ex": (_i, [_p, _l, _i, _p, _i, _p, _p, _p, _p, _l, _p, _p, _p, _i, _l, _p, _p, _p]),
    "map_scatter_rows": (_i, [_p, _p, _p, _l, _i, _p, _p]),
    "map_adamw_hyper_step": (_i, [_p, _p, _d, _d, _d, _d, _i, _l, _l, _p]),
    "map_adamw_hyper_set": (_i, [_p, _d, _d, _d, _d, _l, _p]),
    "map_adamw_multi_tensor": (_i, [_p, _i, _l, _p, _p]),
    "map_adamw_sparse_rows": (_i, [_p, _p, _p, _i, _p, _p, _p, _l, _p, _f, _p]),
    "map_adamw_dense_rows_sparse_grad": (_i, [_p, _p, _p, _l, _i, _p, _p, _p, _p, _f, _p]),
    "map_mask_index_philox": (_i, [_p, _l, _i, _i, _i, _u64, _u64, _l, _p, _p]),
    "map_mfp_mask_apply": (_i, [_p, _p, _l, _i, _i, _l, _p, _p, _p]),
    "map_rfd_replace_philox": (_i, [_p, _p, _l, _i, _i, _i, _p, _l, _p, _p, _l, _u64, _u64, _u64, _l, _p, _p, _p, _p, _p]),
    "map_alias_build": (_i, [_p, _l, _p, _p]),
    "map_alias_draw_philox": (_i, [_p, _p, _l, _u64, _u64, _l, _l, _p, _p, _p]),
    "map_nce_fwd": (_i, [_p, _l, _i, _i, _p, _p, _p, _p, _p, _l, _f, _i, _f, _p, _p, _p, _p, _p, _p, _p]),
    "map_expand_slices": (_i, [_p, _p, _l, _i, _i, _i, _p, _l, _p, _l, _l, _i, _p]),
    "map_nce_full_ce_workspace_bytes": (_sz, [_l, _l]),
    "map_nce_full_ce": (_i, [_p, _l, _i, _p, _p, _l, _p, _p, _p, _sz, _p]),
    "map_gather_slices": (_i, [_p, _p, _l, _i, _i, _i, _p, _p]),
    "map_scatter_add_slices": (_i, [_p, _p, _l, _i, _i, _i, _p, _p]),
    "map_field_enc_supported": (_i, [_i, _i]),
    "map_field_bucket": (_i, [_p, _l, _i, _p, _p, _p]),
    "map_field_enc_fwd": (_i, [_p, _l, _i, _p, _l, _p, _p, _p, _l, _i, _i, _i, _p, _p]),
    "map_field_enc_dgrad": (_i, [_p, _p, _l, _i, _p, _p, _l, _i, _i, _p, _l, _p]),
    "map_field_enc_wgrad": (_i, [_p, _p, _l, _i, _p, _p, _l, _i, _i, _i, _p, _l, _p, _p]),
    "map_head_bwd_fold": (_i, [C.POINTER(HeadBwdArgs), _p]),
    "map_cin_relayout": (_i, [_p, _p, _l, _i, _i, _l, _i, _i, _p]),
    "map_cin_hadamard_fwd": (_i, [_p, _l, _p, _l, _l, _i, _i, _p, _l, _p]),
    "map_cin_hadamard_bwd": (_i, [_p, _l, _p, _l, _p, _l, _l, _i, _i, _p, _l, _i, _p, _l, _p]),
    "map_cin_pool_fwd": (_i, [_p, _l, _l, _i, _i, _p, _l, _p]),
    "map_cin_pool_bwd": (_i, [_p, _l, _p, _l, _l, _i, _i, _p, _l, _p]),
    "map_reduce_sum_f32": (_i, [_p, _l, _f, _p, _p, _sz, _p]),
    "map_reduce_workspace_bytes": (_sz, [_l]),
    "map_bce_logits_fwd": (_i, [_p, _p, _l, _p, _p, _p, _sz, _p]),
    "map_float_sort_keys": (_i, [_p, _l, _p, _p]),
    "map_auc_rank_sum": (_i, [_p, _p, _p, _l, _p, _p]),
    "map_fm_lr_fwd": (_i, [_p, _p, _p, _p, _l, _i, _i, _p, _l, _p]),
    "map_fm_lr_bwd": (_i, [_p, _p, _l, _l, _i, _i, _i, _p, _p, _p]),
    "map_gemm_f32_simt": (_i, [C.POINTER(GemmArgs), _p]),
    "map_gemm_tf32_tcgen05": (_i, [C.POINTER(GemmArgs), _p]),
    "map_gemm_tf32_group": (_i, [C.POINTER(GemmArgs), _i, _p]),
    "map_gemm_tf32_supported": (_i, [C.POINTER(GemmArgs)]),
    "map_gemm_bf16s_group": (_i, [C.POINTER(GemmSplitArgs), _i, _p]),
    "map_gemm_bf16s_set_trace": (_i, [_p, _l]),
    "map_split_bf16": (_i, [_p, _l, _l, _l, _p, _l, _l, _i, _p]),
    "map_gemm_set_trace": (_i, [_p, _l]),
    "map_emb_gather_owned_f32": (_i, [_p, _l, _i, _p, _l, _i, _i, _p, _p]),
    "map_owned_keys": (_i, [_p, _l, _i, _i, _l, _p, _p]),
    "map_nce_scores_owned": (_i, [_p, _l, _i, _i, _p, _p, _p, _i, _i, _p, _p]),
    "map_nce_loss_from_scores": (_i, [_p, _p, _l, _i, _p, _f, _i, _f, _p, _p, _p, _p, _p]),
    "map_nce_dinput_owned": (_i, [_p, _l, _i, _i, _p, _p, _i, _i, _p, _p]),
    "map_p2p_alloc": (_i, [_sz, C.POINTER(C.c_void_p), _p]),
    "map_p2p_open": (_i, [_p, C.POINTER(C.c_void_p)]),
    "map_p2p_close": (_i, [_p]),
    "map_p2p_free": (_i, [_p]),
    "map_emb_gather_sharded_f32": (_i, [_p, _i, _l, _i, _p, _l, _p, _p, _p]),
    "map_nce_fwd_sharded": (_i, [_p, _l, _i, _i, _p, _p, _p, _p, _i, _p, _f, _i, _f, _p, _p, _p, _p, _p, _p, _p]),
    "map_nce_ids_concat": (_i, [_p, _p, _l, _i, _p, _p]),
    "map_p2p_barrier": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "map_owned_compact": (_i, [_p, _p, _i, _i, _l, _p, _p, _p, _p]),
    "map_p2p_reduce_f32": (_i, [_p, _i, _l, _l, _p, _p]),
    "map_p2p_gather_slices_f32": (_i, [_p, _i, _l, _l, _l, _p, _p]),
    "map_colsum_f32": (_i, [_p, _l, _l, _i, _p, _p, _sz, _p]),
    "map_colsum_workspace_bytes": (_sz, [_l, _i]),
    "map_cross_bwd_pre": (_i, [_p, _l, _p, _l, _p, _l, _l, _i, _i, _p, _p, _p]),
    "map_add3_f32": (_i, [_p, _l, _p, _l, _p, _l, _l, _i, _p, _l, _p]),
    "map_relu_bwd_f32": (_i, [_p, _l, _p, _l, _l, _i, _p, _l, _p]),
    "map_scale_by_scalar_f32": (_i, [_p, _p, _l, _p, _p]),
    "map_copy2d_f32": (_i, [_p, _l, _l, _i, _p, _l, _p]),
    "map_gather_rows_i64": (_i, [_p, _l, _i, _p, _l, _p, _p]),
    "map_transpose_f32": (_i, [_p, _l, _l, _l, _p, _l, _p]),
}

_lib = None


class MapB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads the library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MapB200Error(
            f"{LIB_PATH} not found: the sm_100a kernels are not built (run `python -c 'import __graft_entry__ as g; "
            f"g.build()'`).  map_code_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.map_abi_version() != 1:
        raise MapB200Error(f"ABI version mismatch: library {lib.map_abi_version()} != binding 1")
    _lib = lib
    return lib


def last_error() -> str:
    return load().map_last_error().decode("utf-8", "replace")


# number of kernels one entry point launches (for the `gpu_launches` figure of bench.py)
KERNELS_PER_CALL = {
    "map_segment_reduce_rows_ex": 2, "map_p2p_alloc": 0, "map_p2p_open": 0, "map_p2p_close": 0, "map_p2p_free": 0,
    "map_segment_reduce_rows": 2, "map_reduce_sum_f32": 2, "map_bce_logits_fwd": 2, "map_colsum_f32": 2, "map_nce_full_ce": 2,
    "map_alias_build": 0,
}
PROFILE = None        # set to a list: every call is bracketed by CUDA events -> (name, tag, start_event, end_event)
CURRENT_TAG = None    # free-form shape tag set by ops.* just before a call (e.g. "4096x1000x624 tn")
LAUNCHES = None       # set to a dict: name -> number of kernels launched
RECORD = None         # set to a list: (name, args, tag) of every call (arguments kept alive), for replaying one kernel class alone
TIMELINE = None       # set to dict(buf=<device int64 tensor>, ops=[]): a timestamp marker follows every call on its stream
HOST_FUNCS = {"map_alias_build", "map_gemm_set_trace", "map_gemm_bf16s_set_trace", "map_p2p_alloc", "map_p2p_open", "map_p2p_close", "map_p2p_free"}


def mark(name: str, tag=None, stream: int = None):
    """Timeline marker after a non-library operation (e.g. an NCCL collective issued through torch.distributed)."""
    if TIMELINE is None:
        return
    import torch
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    i = len(TIMELINE["ops"])
    if i < TIMELINE["buf"].numel():
        TIMELINE["ops"].append((name, tag, int(st or 0)))
        load().map_timestamp_ns(TIMELINE["buf"].data_ptr() + 8 * i, st)


def call(name: str, *args):
    """Calls an int-returning entry point and raises MapB200Error(map_last_error()) on a non-zero code.
    MAP_EUNSUPPORTED is raised as NotImplementedError (mirrors the reference's NotImplementedError sites)."""
    global CURRENT_TAG
    if LAUNCHES is not None:
        k = KERNELS_PER_CALL.get(name, 1)
        LAUNCHES[name] = LAUNCHES.get(name, 0) + (k(args) if callable(k) else k)
    if RECORD is not None:
        RECORD.append((name, args, CURRENT_TAG))
        if PROFILE is None and TIMELINE is None:
            CURRENT_TAG = None
    if PROFILE is not None:
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = getattr(load(), name)(*args)
        ev1.record()
        PROFILE.append((name, CURRENT_TAG, ev0, ev1))
        CURRENT_TAG = None
    else:
        rc = getattr(load(), name)(*args)
    if TIMELINE is not None and name not in HOST_FUNCS and rc == MAP_OK:
        i = len(TIMELINE["ops"])
        if i < TIMELINE["buf"].numel():
            TIMELINE["ops"].append((name, CURRENT_TAG if PROFILE is None else None, int(args[-1] or 0)))
            load().map_timestamp_ns(TIMELINE["buf"].data_ptr() + 8 * i, args[-1])
        CURRENT_TAG = None
    if rc != MAP_OK:
        msg = f"{name} failed ({rc}): {last_error()}"
        if rc == MAP_EUNSUPPORTED:
            raise NotImplementedError(msg)
        raise MapB200Error(msg)
