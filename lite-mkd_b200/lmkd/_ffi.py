"""ctypes binding of liblmkd.so (include/lmkd.h).

There is no fallback: if the shared library is missing or a tensor is not on a CUDA device the
call raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` from the repo root
(or `make -C lite-mkd_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblmkd.so")
_lib = None

vp, i32, i64, f32, u64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t


class TrxShape(C.Structure):
    _fields_ = [("B", i32), ("Ns", i32), ("Nq", i32), ("L", i32), ("D", i32), ("d", i32), ("card", i32),
                ("way", i32), ("shot", i32), ("dropout_p", f32), ("seed", u64), ("seed_dev", vp), ("ln_eps", f32)]


class LossTerm(C.Structure):
    _fields_ = [("kind", i32), ("rows", i32), ("cols", i32), ("s", vp), ("t", vp), ("y", vp), ("grad", vp),
                ("grad_accumulate", i32), ("w", f32), ("fa", f32), ("fb", f32)]


class FusionLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("w_qkv", "b_qkv", "w_o", "b_o", "w_ff1", "b_ff1", "w_ff2", "b_ff2",
                                  "ln1_g", "ln1_b", "ln2_g", "ln2_b")]


class FusionEncoder(C.Structure):
    _fields_ = [("nmod", i32), ("dmod", i32), ("nhead", i32), ("dff", i32), ("nlayers", i32), ("dout", i32),
                ("ln_eps", f32), ("pe_emb", vp * 4), ("pe_g", vp * 4), ("pe_b", vp * 4),
                ("layers", C.POINTER(FusionLayer)), ("w_out", vp), ("b_out", vp)]


# name -> (restype, argtypes); every symbol include/lmkd.h declares
SIGNATURES = {
    "lmkd_last_error": (C.c_char_p, []),
    "lmkd_version": (i32, []),
    "lmkd_sim_pitch": (i64, [i64]),
    "lmkd_sim_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "lmkd_sim_fwd": (i32, [vp, vp, i32, i32, i32, i32, f32, vp, vp, vp]),
    "lmkd_otam_workspace_bytes": (sz, [i32, i32, i32, i32, i32, i32]),
    "lmkd_otam_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, vp, vp, vp, vp, vp]),
    "lmkd_otam_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, vp, vp, vp, vp]),
    "lmkd_otam_cum_dist": (i32, [vp, i64, i32, i32, f32, vp, vp, vp, vp]),
    "lmkd_trx_workspace_bytes": (sz, [C.POINTER(TrxShape), i32]),
    "lmkd_trx_fwd": (i32, [C.POINTER(TrxShape), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]),
    "lmkd_trx_bwd": (i32, [C.POINTER(TrxShape), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "lmkd_trx_attn_fused_fits": (i32, [C.POINTER(TrxShape)]),
    "lmkd_trx_set_attn_budget": (None, [C.c_double]),
    "lmkd_trx_attn_fwd": (i32, [C.POINTER(TrxShape), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "lmkd_strm_dist_workspace_bytes": (sz, [C.POINTER(TrxShape), i32]),
    "lmkd_strm_dist_fwd": (i32, [C.POINTER(TrxShape), vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]),
    "lmkd_strm_dist_bwd": (i32, [C.POINTER(TrxShape), vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "lmkd_dropout_mask": (i32, [vp, i64, f32, u64, vp]),
    "lmkd_support_dk_fwd": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "lmkd_support_dk_bwd": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "lmkd_edist_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "lmkd_edist_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "lmkd_edist_bwd": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "lmkd_frame_pool_fwd": (i32, [vp, i64, i32, i32, i32, i32, vp, vp]),
    "lmkd_frame_pool_bwd": (i32, [vp, vp, i64, i32, i32, i32, i32, vp, vp]),
    "lmkd_feature_head_workspace_bytes": (sz, [i64, i32, i32, i32]),
    "lmkd_feature_head_fwd": (i32, [vp, vp, vp, i64, i32, i32, i32, vp, vp, vp]),
    "lmkd_feature_head_bwd": (i32, [vp, i64, i32, i32, i32, vp, vp, vp, vp, vp]),
    "lmkd_d2m_logit_loss": (i32, [C.POINTER(LossTerm), i32, f32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "lmkd_mse_partials": (i32, []),
    "lmkd_d2m_feature_mse_fwdbwd": (i32, [vp, vp, vp, i64, i32, f32, f32, vp, vp, i32, vp]),
    "lmkd_episode_gather": (i32, [vp, i32, i64, vp, i64, i64, vp, vp, vp]),
    "lmkd_d2m_feature_mse_store_fwdbwd": (i32, [vp, vp, i32, i64, vp, i64, i64, vp, f32, f32, vp, vp, i32, vp, vp]),
    "lmkd_fusion_workspace_bytes": (sz, [C.POINTER(FusionEncoder), i64, i32]),
    "lmkd_fusion_fwd": (i32, [C.POINTER(FusionEncoder), C.POINTER(vp), C.POINTER(i32), i64, i32, vp, i32, vp, vp]),
    "lmkd_scale_by_device_scalar": (i32, [vp, i64, vp, vp]),
    "lmkd_accuracy_count": (i32, [vp, vp, i64, i32, vp, vp]),
    "lmkd_gemm_bf16": (i32, [i32, i32, i32, i32, vp, i32, i64, i64, vp, i32, i64, i64, vp, i64, i64, f32, i32, i32, vp]),
    "lmkd_cast_bf16": (i32, [vp, vp, i64, vp]),
    "lmkd_upcast_bf16": (i32, [vp, vp, i64, vp]),
    "lmkd_launch_count": (C.c_longlong, [i32]),
    "lmkd_gemm_timing_enable": (None, [i32]),
    "lmkd_gemm_timing_read": (i32, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i32)]),
    "lmkd_kernel_timing_read": (i32, [i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i32)]),
}


def lib():
    """Load liblmkd.so once; raise (never fall back) if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` in the repo root. "
                "There is no CPU or PyTorch fallback for this path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "lmkd") -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().lmkd_last_error().decode()}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("lmkd operates on CUDA tensors only (no CPU fallback); got a CPU tensor")
    if not t.is_contiguous():
        raise RuntimeError("lmkd needs contiguous tensors")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous view of a CUDA tensor (copies only when needed)."""
    if not t.is_cuda:
        raise RuntimeError("lmkd operates on CUDA tensors only (no CPU fallback); got a CPU tensor")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_status = {}


class _Status:
    """One int32 in PINNED HOST memory per device.  Under unified addressing the kernels OR their error bits
    straight into it (a store over PCIe, only ever executed on an error), and the host can look at it without
    synchronising: every wrapper polls it on entry, so a bad label or index raises at the latest one call after
    the kernel that saw it ran -- no product path has to remember to ask."""

    def __init__(self):
        self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.ptr = C.c_void_p(self.host.data_ptr())

    def take(self) -> int:
        v = int(self.host[0])
        if v:
            self.host[0] = 0
        return v


def status_word(device) -> _Status:
    key = torch.device(device).index or 0
    if key not in _status:
        _status[key] = _Status()
    return _status[key]


def status_ptr(device):
    """Pointer the kernels get as `int* status` (pinned host memory, device-accessible)."""
    return status_word(device).ptr


def _raise_status(v: int) -> None:
    msgs = []
    if v & 1:
        msgs.append("a support label lies outside [0, way)")
    if v & 2:
        msgs.append("a class has more supports than `shot`")
    if v & 4:
        msgs.append("a feature-store index lies outside the store")
    raise RuntimeError("lmkd: " + "; ".join(msgs) + " (reported by an earlier kernel launch)")


def poll_status(device=None) -> None:
    """Non-synchronising look at the error bits; called on entry by every op wrapper."""
    for key, st in _status.items():
        if device is not None and (torch.device(device).index or 0) != key:
            continue
        v = st.take()
        if v:
            _raise_status(v)


def check_device_status(device=None) -> None:
    """Synchronising check of the error bits (call at a natural sync point, e.g. where the loss is read)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize(device)
    poll_status(device)
