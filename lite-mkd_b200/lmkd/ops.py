"""torch.autograd wrappers over the C-ABI (one Function per fwd/bwd pair of include/lmkd.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; every arithmetic op of
the hot path runs inside liblmkd.so.  All tensors must be CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import math
from itertools import combinations

import torch

from . import _ffi
from ._ffi import LossTerm, TrxShape, check, f32c, lib, ptr, stream

TERM_CE, TERM_KD, TERM_ICR = 0, 1, 2


def _bytes(n: int, device) -> torch.Tensor:
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------
# OTAM
# --------------------------------------------------------------------------------------------
class _OtamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support, labels, query, way, lbda, eps):
        B, Ns, L, D = support.shape
        Nq = query.shape[1]
        dev = support.device
        ws = _bytes(lib().lmkd_otam_workspace_bytes(B, Ns, Nq, L, D, way), dev)
        probs = torch.empty(B, Nq, way, dtype=torch.float32, device=dev)
        check(lib().lmkd_otam_fwd(ptr(support), ptr(labels), ptr(query), B, Ns, Nq, L, D, way, lbda, eps, ptr(probs),
                                  None, ptr(ws), _ffi.status_ptr(dev), stream()), "lmkd_otam_fwd")
        ctx.save_for_backward(support, labels, query, probs, ws)
        ctx.cfg = (way, lbda, eps)
        return probs

    @staticmethod
    def backward(ctx, gprobs):
        support, labels, query, probs, ws = ctx.saved_tensors
        way, lbda, eps = ctx.cfg
        B, Ns, L, D = support.shape
        Nq = query.shape[1]
        gs, gq = torch.empty_like(support), torch.empty_like(query)
        check(lib().lmkd_otam_bwd(ptr(f32c(gprobs)), ptr(probs), ptr(support), ptr(labels), ptr(query), B, Ns, Nq, L, D,
                                  way, lbda, eps, ptr(gs), ptr(gq), ptr(ws), stream()), "lmkd_otam_bwd")
        return gs, None, gq, None, None, None


def otam_probs(support, labels, query, way: int, lbda: float = 0.1, eps: float = 0.01):
    """[B,Ns,L,D], [B,Ns], [B,Nq,L,D] -> [B,Nq,way] (CNN_OTAM.forward, teacher/code/model.py:3319-3343)."""
    _ffi.poll_status(support.device)
    return _OtamFn.apply(f32c(support), f32c(labels), f32c(query), int(way), float(lbda), float(eps))


def otam_cum_dist(dists, lbda: float = 0.1, grad_out=None):
    """OTAM_cum_dist on [..., L, M] (one direction); returns (out, grad_dists or None)."""
    d = f32c(dists)
    lead = d.shape[:-2]
    L, M = d.shape[-2:]
    P = int(math.prod(lead)) if lead else 1
    out = torch.empty(P, dtype=torch.float32, device=d.device)
    gd = torch.empty_like(d) if grad_out is not None else None
    go = f32c(grad_out).reshape(P) if grad_out is not None else None
    check(lib().lmkd_otam_cum_dist(ptr(d), P, L, M, lbda, ptr(out), ptr(go), ptr(gd), stream()), "lmkd_otam_cum_dist")
    return out.reshape(lead), gd


def frame_dists(x, y, eps: float = 0.01):
    """1 - cos_sim(x, y) (teacher/code/model.py:3260-3269,3333): [B,nx,D],[B,ny,D] -> [B,nx,ny]."""
    x, y = f32c(x), f32c(y)
    B, nx, D = x.shape
    ny = y.shape[1]
    ld = lib().lmkd_sim_pitch(ny)
    ws = _bytes(lib().lmkd_sim_workspace_bytes(B, nx, ny, D), x.device)
    dist = torch.empty(B, nx, ld, dtype=torch.float32, device=x.device)
    check(lib().lmkd_sim_fwd(ptr(x), ptr(y), B, nx, ny, D, eps, ptr(dist), ptr(ws), stream()), "lmkd_sim_fwd")
    return dist[:, :, :ny]


# --------------------------------------------------------------------------------------------
# TRX
# --------------------------------------------------------------------------------------------
def tuple_tables(seq_len: int, card: int):
    """(tuples [T,c], inv_off [c*L+1], inv_idx [c*T]) as CPU int32 tensors (TRX.py:70-73)."""
    tuples = list(combinations(range(seq_len), card))
    inv_off, inv_idx = [0], []
    for j in range(card):
        for l in range(seq_len):
            inv_idx.extend(t for t, tp in enumerate(tuples) if tp[j] == l)
            inv_off.append(len(inv_idx))
    return (torch.tensor(tuples, dtype=torch.int32).reshape(len(tuples), card),
            torch.tensor(inv_off, dtype=torch.int32), torch.tensor(inv_idx, dtype=torch.int32))


# Opt-in (bench.py, training loops that own their gradient buffers): when every head parameter already has a
# contiguous fp32 .grad, the TRX backward ADDS its parameter gradients into those buffers inside its own kernels and
# returns None for them, instead of handing fresh tensors to autograd for a separate accumulate pass (12 elementwise
# kernels moving 3 x 94 MB per micro-batch at config 2).  Parameter hooks do not fire in this mode.
ACCUMULATE_PARAM_GRADS_IN_PLACE = False


def _trx_forward(support, labels, query, pe, params, tables, cfg, need_grad):
    """lmkd_trx_fwd for one cardinality; returns (logits, sim, saved state for _trx_backward or None)."""
    Wk, bk, Wv, bv, gamma, beta = params
    tuples, inv_off, inv_idx = tables
    B, Ns, L, D = support.shape
    Nq = query.shape[1]
    d, card, way, shot, p, seed, ln_eps, with_sim, seed_dev = cfg
    shape = TrxShape(B, Ns, Nq, L, D, d, card, way, shot, p, seed, ptr(seed_dev), ln_eps)
    need = int(need_grad) * (2 if with_sim else 1)
    dev = support.device
    nbytes = lib().lmkd_trx_workspace_bytes(C.byref(shape), need)
    if nbytes == 0:
        raise RuntimeError("lmkd_trx_workspace_bytes: " + lib().lmkd_last_error().decode())
    ws = _bytes(nbytes, dev)
    logits = torch.empty(B, Nq, way, dtype=torch.float32, device=dev)
    sim = torch.empty(B, Nq, way, way, dtype=torch.float32, device=dev) if with_sim else None
    check(lib().lmkd_trx_fwd(C.byref(shape), ptr(support), ptr(labels), ptr(query), ptr(pe), ptr(tuples), ptr(Wk),
                             ptr(bk), ptr(Wv), ptr(bv), ptr(gamma), ptr(beta), ptr(logits), ptr(sim), ptr(ws),
                             need, _ffi.status_ptr(dev), stream()), "lmkd_trx_fwd")
    state = (ws, shape, need, with_sim) if need else None
    return logits, sim, state


def _trx_backward(state, tables, params, glogits, gsim, gs, gq, accumulate_features):
    """lmkd_trx_bwd for one cardinality into gs / gq (added to when `accumulate_features`); returns the six
    parameter gradients, or None when they were added straight into the parameters' .grad buffers."""
    ws, shape, need, with_sim = state
    tuples, inv_off, inv_idx = tables
    Wk, bk, Wv, bv, gamma, beta = params
    dev = ws.device
    inplace = ACCUMULATE_PARAM_GRADS_IN_PLACE and all(
        p.requires_grad and p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32
        and p.grad.shape == p.shape for p in params)
    if inplace:
        gWk, gbk, gWv, gbv, gg, gb = (p.grad for p in params)
    else:
        gWk, gWv = torch.empty_like(Wk), torch.empty_like(Wk)
        gbk, gbv, gg, gb = (torch.empty_like(bk) for _ in range(4))
    if glogits is None:
        glogits = torch.zeros(shape.B, shape.Nq, shape.way, dtype=torch.float32, device=dev)
    gsim_c = f32c(gsim) if (with_sim and gsim is not None) else None
    flags = int(inplace) | (2 if accumulate_features else 0)
    check(lib().lmkd_trx_bwd(C.byref(shape), ptr(f32c(glogits)), ptr(gsim_c), ptr(tuples), ptr(inv_off), ptr(inv_idx),
                             ptr(bk), ptr(gamma), ptr(beta), ptr(gs), ptr(gq), ptr(gWk), ptr(gbk), ptr(gWv), ptr(gbv),
                             ptr(gg), ptr(gb), ptr(ws), need, flags, stream()), "lmkd_trx_bwd")
    return None if inplace else (gWk, gbk, gWv, gbv, gg, gb)


class _TrxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support, labels, query, pe, Wk, bk, Wv, bv, gamma, beta, tables, cfg):
        params = (Wk, bk, Wv, bv, gamma, beta)
        logits, sim, state = _trx_forward(support, labels, query, pe, params, tables, cfg, any(ctx.needs_input_grad))
        ctx.state, ctx.tables, ctx.params = state, tables, params
        ctx.sizes = (support.shape, query.shape)
        ctx.with_sim = cfg[7]
        return (logits, sim) if ctx.with_sim else logits

    @staticmethod
    def backward(ctx, glogits, gsim=None):
        dev = ctx.state[0].device
        gs = torch.empty(ctx.sizes[0], dtype=torch.float32, device=dev)
        gq = torch.empty(ctx.sizes[1], dtype=torch.float32, device=dev)
        g = _trx_backward(ctx.state, ctx.tables, ctx.params, glogits, gsim, gs, gq, False)
        if g is None:
            return gs, None, gq, None, None, None, None, None, None, None, None, None
        return (gs, None, gq, None) + g + (None, None)


class _TrxBranchFn(torch.autograd.Function):
    """TrxBranch (teacher/code/model.py:1094-1128): every cardinality on the same episode, logits averaged.  One
    autograd node for the whole branch, so the feature gradients of the cardinalities are summed inside the
    backward kernels (trx_dx_scatter adds into the buffer the first cardinality wrote) instead of by separate
    elementwise passes over two 105 MB tensors per side at config 2."""

    @staticmethod
    def forward(ctx, support, labels, query, heads, *flat):
        # heads: list of (pe, tables, cfg); flat: 6 parameters per cardinality
        n = len(heads)
        need = any(ctx.needs_input_grad)
        states, total = [], None
        for i, (pe, tables, cfg) in enumerate(heads):
            logits, _, st = _trx_forward(support, labels, query, pe, flat[6 * i:6 * i + 6], tables, cfg, need)
            states.append(st)
            total = logits if total is None else total.add_(logits)
        if n > 1:
            total.mul_(1.0 / n)
        ctx.states, ctx.heads, ctx.flat = states, heads, flat
        ctx.sizes = (support.shape, query.shape)
        return total

    @staticmethod
    def backward(ctx, glogits):
        n = len(ctx.heads)
        dev = ctx.states[0][0].device
        gs = torch.empty(ctx.sizes[0], dtype=torch.float32, device=dev)
        gq = torch.empty(ctx.sizes[1], dtype=torch.float32, device=dev)
        g_each = f32c(glogits) if n == 1 else f32c(glogits) * (1.0 / n)
        grads = []
        for i, (pe, tables, cfg) in enumerate(ctx.heads):
            g = _trx_backward(ctx.states[i], tables, ctx.flat[6 * i:6 * i + 6], g_each, None, gs, gq, i > 0)
            grads.extend(g if g is not None else (None,) * 6)
        return (gs, None, gq, None) + tuple(grads)


def trx_branch_logits(support, labels, query, heads):
    """heads: list of dicts with keys pe, Wk, bk, Wv, bv, gamma, beta, tables, card, way, shot, dropout_p, seed,
    ln_eps, seed_dev (the arguments of trx_logits, one entry per cardinality) -> mean logits [B, Nq, way]."""
    _ffi.poll_status(support.device)
    packed, flat = [], []
    for h in heads:
        cfg = (int(h["Wk"].shape[0]), int(h["card"]), int(h["way"]), int(h["shot"]), float(h.get("dropout_p", 0.0)),
               int(h.get("seed", 0)), float(h.get("ln_eps", 1e-5)), False, h.get("seed_dev"))
        packed.append((f32c(h["pe"]), h["tables"], cfg))
        flat.extend(f32c(h[k]) for k in ("Wk", "bk", "Wv", "bv", "gamma", "beta"))
    return _TrxBranchFn.apply(f32c(support), f32c(labels), f32c(query), packed, *flat)


def trx_logits(support, labels, query, pe, Wk, bk, Wv, bv, gamma, beta, tables, *, card, way, shot,
               dropout_p=0.0, seed=0, ln_eps=1e-5, with_proto_sim=False, seed_dev=None):
    """One-cardinality TemporalCrossTransformer on batched episodes -> logits [B, Nq, way]
    (and, with_proto_sim, the TRX_sup prototype cosine matrix [B, Nq, way, way])."""
    _ffi.poll_status(support.device)
    cfg = (int(Wk.shape[0]), int(card), int(way), int(shot), float(dropout_p), int(seed), float(ln_eps),
           bool(with_proto_sim), seed_dev)
    return _TrxFn.apply(f32c(support), f32c(labels), f32c(query), f32c(pe), f32c(Wk), f32c(bk), f32c(Wv), f32c(bv),
                        f32c(gamma), f32c(beta), tables, cfg)


class _StrmDistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support, labels, query, W, bias, tables, cfg):
        tuples, inv_off, inv_idx = tables
        B, Ns, L, D = support.shape
        Nq = query.shape[1]
        card, way, shot, p, seed, seed_dev = cfg
        shape = TrxShape(B, Ns, Nq, L, D, int(W.shape[0]), card, way, shot, p, seed, ptr(seed_dev), 1e-5)
        need_grad = int(any(ctx.needs_input_grad))
        dev = support.device
        nbytes = lib().lmkd_strm_dist_workspace_bytes(C.byref(shape), need_grad)
        if nbytes == 0:
            raise RuntimeError("lmkd_strm_dist_workspace_bytes: " + lib().lmkd_last_error().decode())
        ws = _bytes(nbytes, dev)
        logits = torch.empty(B, Nq, way, dtype=torch.float32, device=dev)
        check(lib().lmkd_strm_dist_fwd(C.byref(shape), ptr(support), ptr(labels), ptr(query), ptr(tuples), ptr(W), ptr(bias),
                                       ptr(logits), ptr(ws), need_grad, _ffi.status_ptr(dev), stream()), "lmkd_strm_dist_fwd")
        if need_grad:
            ctx.save_for_backward(ws, inv_off, inv_idx, W, bias)
            ctx.shape = shape
            ctx.sizes = (support.shape, query.shape)
        return logits

    @staticmethod
    def backward(ctx, glogits):
        ws, inv_off, inv_idx, W, bias = ctx.saved_tensors
        dev = ws.device
        gs = torch.empty(ctx.sizes[0], dtype=torch.float32, device=dev)
        gq = torch.empty(ctx.sizes[1], dtype=torch.float32, device=dev)
        gW, gb = torch.empty_like(W), torch.empty_like(bias)
        check(lib().lmkd_strm_dist_bwd(C.byref(ctx.shape), ptr(f32c(glogits)), ptr(inv_off), ptr(inv_idx), ptr(gs), ptr(gq),
                                       ptr(gW), ptr(gb), ptr(ws), stream()), "lmkd_strm_dist_bwd")
        return gs, None, gq, gW, gb, None, None


def strm_distance_logits(support, labels, query, W, bias, tables, *, card, way, shot, dropout_p=0.0, seed=0,
                         seed_dev=None):
    """STRM DistanceLoss on batched episodes -> logits [B, Nq, way] (strm_res18_sup.py:184-243)."""
    _ffi.poll_status(support.device)
    cfg = (int(card), int(way), int(shot), float(dropout_p), int(seed), seed_dev)
    return _StrmDistFn.apply(f32c(support), f32c(labels), f32c(query), f32c(W), f32c(bias), tables, cfg)


def dropout_mask(n: int, p: float, seed: int, device) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(lib().lmkd_dropout_mask(ptr(out), n, p, seed, stream()), "lmkd_dropout_mask")
    return out


# --------------------------------------------------------------------------------------------
# frame-mean Euclidean heads (e_dist / CosDistance)
# --------------------------------------------------------------------------------------------
class _EdistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support, labels, query, way):
        B, Ns, L, D = support.shape
        Nq = query.shape[1]
        dev = support.device
        ws = _bytes(lib().lmkd_edist_workspace_bytes(B, Ns, Nq, D), dev)
        logits = torch.empty(B, Nq, way, dtype=torch.float32, device=dev)
        check(lib().lmkd_edist_fwd(ptr(support), ptr(labels), ptr(query), B, Ns, Nq, L, D, way, ptr(logits), ptr(ws),
                                   _ffi.status_ptr(dev), stream()), "lmkd_edist_fwd")
        ctx.save_for_backward(labels, ws)
        ctx.cfg = (B, Ns, Nq, L, D, way)
        return logits

    @staticmethod
    def backward(ctx, glogits):
        labels, ws = ctx.saved_tensors
        B, Ns, Nq, L, D, way = ctx.cfg
        gs = torch.empty(B, Ns, L, D, dtype=torch.float32, device=ws.device)
        gq = torch.empty(B, Nq, L, D, dtype=torch.float32, device=ws.device)
        check(lib().lmkd_edist_bwd(ptr(f32c(glogits)), ptr(labels), B, Ns, Nq, L, D, way, ptr(gs), ptr(gq), ptr(ws),
                                   stream()), "lmkd_edist_bwd")
        return gs, None, gq, None


def edist_logits(support, labels, query, way: int):
    """[B,Ns,L,D], [B,Ns], [B,Nq,L,D] -> [B,Nq,way] (e_dist.py:22-61)."""
    _ffi.poll_status(support.device)
    return _EdistFn.apply(f32c(support), f32c(labels), f32c(query), int(way))


# --------------------------------------------------------------------------------------------
# Student feature heads feeding the path
# --------------------------------------------------------------------------------------------
class _FramePoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap, out_hw):
        rows, Cc, H, W = fmap.shape
        pooled = torch.empty(rows, Cc, dtype=torch.float32, device=fmap.device)
        check(lib().lmkd_frame_pool_fwd(ptr(fmap), rows, Cc, H, W, out_hw, ptr(pooled), stream()), "lmkd_frame_pool_fwd")
        ctx.save_for_backward(fmap)
        ctx.out_hw = out_hw
        return pooled

    @staticmethod
    def backward(ctx, gpooled):
        (fmap,) = ctx.saved_tensors
        rows, Cc, H, W = fmap.shape
        gmap = torch.empty_like(fmap)
        check(lib().lmkd_frame_pool_bwd(ptr(fmap), ptr(f32c(gpooled)), rows, Cc, H, W, ctx.out_hw, ptr(gmap), stream()),
              "lmkd_frame_pool_bwd")
        return gmap, None


def frame_pool(fmap, out_hw: int = 4):
    """[rows, C, H, W] trunk maps -> [rows, C]: AdaptiveMaxPool2d(out_hw) then the mean over the patches
    (resnet18_2fc.py:41-53)."""
    return _FramePoolFn.apply(f32c(fmap), int(out_hw))


class _FeatureHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        rows, in_dim = x.shape
        heads, out_dim, _ = weight.shape
        dev = x.device
        ws = _bytes(lib().lmkd_feature_head_workspace_bytes(rows, in_dim, out_dim, heads), dev)
        y = torch.empty(heads, rows, out_dim, dtype=torch.float32, device=dev)
        check(lib().lmkd_feature_head_fwd(ptr(x), ptr(weight), ptr(bias), rows, in_dim, out_dim, heads, ptr(y), ptr(ws),
                                          stream()), "lmkd_feature_head_fwd")
        ctx.save_for_backward(ws)
        ctx.cfg = (rows, in_dim, out_dim, heads)
        return y

    @staticmethod
    def backward(ctx, gy):
        (ws,) = ctx.saved_tensors
        rows, in_dim, out_dim, heads = ctx.cfg
        dev = ws.device
        need_x, need_w, need_b = ctx.needs_input_grad
        gx = torch.empty(rows, in_dim, dtype=torch.float32, device=dev) if need_x else None
        gw = torch.empty(heads, out_dim, in_dim, dtype=torch.float32, device=dev) if need_w else None
        gb = torch.empty(heads, out_dim, dtype=torch.float32, device=dev) if need_b else None
        check(lib().lmkd_feature_head_bwd(ptr(f32c(gy)), rows, in_dim, out_dim, heads, ptr(gx) if need_x else None,
                                          ptr(gw) if need_w else None, ptr(gb) if need_b else None, ptr(ws), stream()),
              "lmkd_feature_head_bwd")
        return gx, gw, gb


def feature_heads(x, weight, bias):
    """x [rows, in], weight [heads, out, in], bias [heads, out] -> [heads, rows, out]
    (fc1 / fc2 of resnet18_2fc.py:55-64, res18_2048 of resnet18_student.py:53-54)."""
    return _FeatureHeadFn.apply(f32c(x), f32c(weight), f32c(bias))


# --------------------------------------------------------------------------------------------
# SupportDK
# --------------------------------------------------------------------------------------------
class _SupportDkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, support, way, shot):
        B, Ns, L, D = support.shape
        protos = torch.empty(B, way, L, D, dtype=torch.float32, device=support.device)
        out = torch.empty(B, way, way - 1, dtype=torch.float32, device=support.device)
        check(lib().lmkd_support_dk_fwd(ptr(support), B, way, shot, L, D, ptr(protos), ptr(out), stream()),
              "lmkd_support_dk_fwd")
        ctx.save_for_backward(protos)
        ctx.cfg = (B, way, shot, L, D)
        return out

    @staticmethod
    def backward(ctx, gout):
        (protos,) = ctx.saved_tensors
        B, way, shot, L, D = ctx.cfg
        gs = torch.empty(B, way * shot, L, D, dtype=torch.float32, device=protos.device)
        check(lib().lmkd_support_dk_bwd(ptr(f32c(gout)), ptr(protos), B, way, shot, L, D, ptr(gs), stream()),
              "lmkd_support_dk_bwd")
        return gs, None, None


def support_dk(support, way: int, shot: int):
    """[B, way*shot, L, D] -> [B, way, way-1] (SupportDK, TRX_2fcsup.py:162-189)."""
    if support.shape[1] != way * shot:
        raise RuntimeError(f"SupportDK expects way*shot = {way * shot} supports, got {support.shape[1]}")
    return _SupportDkFn.apply(f32c(support), int(way), int(shot))


# --------------------------------------------------------------------------------------------
# D2M losses
# --------------------------------------------------------------------------------------------
class _LogitLossFn(torch.autograd.Function):
    """All logits-level terms of one recipe in a single launch; forward also produces the gradients."""

    @staticmethod
    def forward(ctx, spec, *students):
        # spec: dict(terms=[(kind, s_idx, target, w, fa, fb)], T, focal=(num_idx, den_idx, labels) or None, B)
        dev = students[0].device
        B = spec["B"]
        grads = [torch.empty_like(s) if ctx.needs_input_grad[i + 1] else None for i, s in enumerate(students)]
        seen = set()
        arr = (LossTerm * len(spec["terms"]))()
        keep = []
        for i, (kind, si, target, w, fa, fb) in enumerate(spec["terms"]):
            s = students[si]
            rows, cols = s.shape[-2], s.shape[-1]
            t_ptr = y_ptr = None
            if kind == TERM_CE:
                y = target.to(device=dev, dtype=torch.int64).contiguous()
                keep.append(y)
                y_ptr = ptr(y)
            else:
                t = f32c(target.detach().to(dev))
                keep.append(t)
                t_ptr = ptr(t)
            arr[i] = LossTerm(kind, rows, cols, ptr(s), t_ptr, y_ptr, ptr(grads[si]), int(si in seen), w, fa, fb)
            if grads[si] is not None:
                seen.add(si)
        for si, g in enumerate(grads):
            if g is not None and si not in seen:
                g.zero_()
        fnum = fden = fy = None
        frows = fcols = 0
        if spec.get("focal") is not None:
            ni, di, lab = spec["focal"]
            fnum, fden = students[ni], students[di]
            fy = lab.to(device=dev, dtype=torch.int64).contiguous()
            frows, fcols = fnum.shape[-2], fnum.shape[-1]
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        values = torch.empty(B, len(spec["terms"]), dtype=torch.float32, device=dev)
        focal = torch.empty(B, dtype=torch.float32, device=dev)
        check(lib().lmkd_d2m_logit_loss(arr, len(spec["terms"]), float(spec["T"]), ptr(fnum), ptr(fden), ptr(fy), frows,
                                        fcols, B, ptr(loss), ptr(values), ptr(focal), stream()), "lmkd_d2m_logit_loss")
        ctx.save_for_backward(*[g for g in grads if g is not None])
        ctx.has_grad = [g is not None for g in grads]
        ctx.mark_non_differentiable(values, focal)
        return loss, values, focal

    @staticmethod
    def backward(ctx, gloss, _gv, _gf):
        saved = list(ctx.saved_tensors)
        out = []
        g = f32c(gloss)
        for has in ctx.has_grad:
            if not has:
                out.append(None)
                continue
            gr = saved.pop(0)
            # gr holds d loss[b] / d s[b]; scale each episode by its upstream gradient
            out.append(gr * g.reshape(-1, *([1] * (gr.dim() - 1))))
        return (None, *out)


def logit_loss(spec, students):
    """Per-episode loss [B], unweighted term values [B, nterms], focal weights [B]."""
    students = [f32c(s) for s in students]
    return _LogitLossFn.apply(spec, *students)


class _FeatureMseFn(torch.autograd.Function):
    """mean((s - t)^2) per episode, summed over episodes; fwd+bwd in one HBM pass."""

    @staticmethod
    def forward(ctx, s, t, weight, n_per_episode):
        dev = s.device
        n = s.numel()
        dtype = {torch.float32: 0, torch.bfloat16: 1}[s.dtype]
        ds = torch.empty_like(s)
        partials = torch.empty(lib().lmkd_mse_partials(), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        check(lib().lmkd_d2m_feature_mse_fwdbwd(ptr(s), ptr(t), ptr(ds), n, dtype, weight / n_per_episode,
                                                2.0 * weight / n_per_episode, ptr(partials), ptr(loss), 0, stream()),
              "lmkd_d2m_feature_mse_fwdbwd")
        ctx.save_for_backward(ds)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (ds,) = ctx.saved_tensors
        g = f32c(g).reshape(1)
        if ds.dtype == torch.float32:
            # The gradient was produced by the forward pass; the upstream scalar is applied IN PLACE by a kernel that
            # exits immediately when it is 1 (the usual case), and the buffer itself is returned.  A second backward
            # through the same node would scale it twice, so it is refused instead of silently returning g^2 * ds.
            if getattr(ctx, "consumed", False):
                raise RuntimeError("lmkd feature-MSE: backward called twice on the same graph (retain_graph); the "
                                   "fused kernel hands out its gradient buffer once -- rebuild the loss instead")
            ctx.consumed = True
            check(lib().lmkd_scale_by_device_scalar(ptr(ds), ds.numel(), ptr(g), stream()), "lmkd_scale")
            return ds, None, None, None
        return ds * g.to(ds.dtype), None, None, None


def feature_mse(student_feature, teacher_feature, weight: float = 1.0, n_per_episode: int | None = None):
    """weight * sum_b mse(s_b, t_b); n_per_episode defaults to the whole tensor (one episode)."""
    s, t = student_feature, teacher_feature
    if not s.is_cuda or not t.is_cuda:
        raise RuntimeError("lmkd operates on CUDA tensors only (no CPU fallback)")
    if s.dtype not in (torch.float32, torch.bfloat16):
        s = s.float()
    t = t.detach().to(s.dtype)
    s, t = s.contiguous(), t.contiguous()
    if s.shape != t.shape:
        raise RuntimeError(f"feature shapes differ: {tuple(s.shape)} vs {tuple(t.shape)}")
    return _FeatureMseFn.apply(s, t, float(weight), int(n_per_episode or s.numel()))


# --------------------------------------------------------------------------------------------
# Teacher-feature store
# --------------------------------------------------------------------------------------------
def _store_args(store, index):
    if not store.is_cuda:
        raise RuntimeError("lmkd operates on CUDA tensors only (no CPU fallback); the feature store is on the CPU")
    if store.dim() != 2 or store.dtype not in (torch.float32, torch.bfloat16) or not store.is_contiguous():
        raise RuntimeError("feature store must be a contiguous [videos, L*D] fp32 or bf16 tensor")
    idx = index.to(device=store.device, dtype=torch.int64).contiguous()
    return idx, {torch.float32: 0, torch.bfloat16: 1}[store.dtype]


def episode_gather(store, index, seq_len: int, out=None):
    """store [videos, L*D] (fp32 / bf16, resident in HBM), index [...] video rows -> [..., L, D] fp32:
    what video_reader.py:388-395 + :470-471 assemble with one np.load per video."""
    _ffi.poll_status(store.device)
    idx, dt = _store_args(store, index)
    row = store.shape[1]
    if out is None:
        out = torch.empty(idx.numel(), row, dtype=torch.float32, device=store.device)
    elif out.dtype != torch.float32 or out.numel() != idx.numel() * row or not out.is_contiguous():
        raise RuntimeError("episode_gather: `out` must be a contiguous fp32 tensor of index.numel() x L*D elements")
    check(lib().lmkd_episode_gather(ptr(store), dt, store.shape[0], ptr(idx), idx.numel(), row, ptr(out),
                                    _ffi.status_ptr(store.device), stream()), "lmkd_episode_gather")
    return out.reshape(*index.shape, seq_len, row // seq_len)


class _FeatureMseStoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, store, idx, dt, weight, n_per_episode):
        dev = s.device
        ds = torch.empty_like(s)
        partials = torch.empty(lib().lmkd_mse_partials(), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        check(lib().lmkd_d2m_feature_mse_store_fwdbwd(ptr(s), ptr(store), dt, store.shape[0], ptr(idx), idx.numel(),
                                                      store.shape[1], ptr(ds), weight / n_per_episode,
                                                      2.0 * weight / n_per_episode, ptr(partials), ptr(loss), 0,
                                                      _ffi.status_ptr(dev), stream()),
              "lmkd_d2m_feature_mse_store_fwdbwd")
        ctx.save_for_backward(ds)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (ds,) = ctx.saved_tensors
        g = f32c(g).reshape(1)
        if getattr(ctx, "consumed", False):     # same one-shot contract as _FeatureMseFn.backward
            raise RuntimeError("lmkd feature-MSE: backward called twice on the same graph (retain_graph); the "
                               "fused kernel hands out its gradient buffer once -- rebuild the loss instead")
        ctx.consumed = True
        check(lib().lmkd_scale_by_device_scalar(ptr(ds), ds.numel(), ptr(g), stream()), "lmkd_scale")
        return ds, None, None, None, None, None


def feature_mse_from_store(student_feature, store, index, weight: float = 1.0, n_per_episode: int | None = None):
    """weight * sum_b mse(s_b, store[index_b]) with the teacher features read in place from the device store:
    student [..., L, D] fp32, index [...] (one store row per video)."""
    _ffi.poll_status(store.device)
    s = f32c(student_feature)
    idx, dt = _store_args(store, index)
    if s.numel() != idx.numel() * store.shape[1]:
        raise RuntimeError(f"student features {tuple(s.shape)} do not match {idx.numel()} store rows of {store.shape[1]}")
    return _FeatureMseStoreFn.apply(s, store, idx, dt, float(weight), int(n_per_episode or s.numel()))


# --------------------------------------------------------------------------------------------
# teacher multi-modal fusion forward
# --------------------------------------------------------------------------------------------
class PackedFusionEncoder:
    """Device-side parameter pack of one fusion encoder (ThreeTransforTemproal / TwoTransforFusion,
    teacher/code/model.py:1300-1392): the Linear matrices cast to bf16 once (lmkd_cast_bf16), everything else kept
    as the module's own fp32 tensors, laid out as the lmkd_fusion_encoder struct of include/lmkd.h.  Rebuilt by the
    owner when a parameter changes (`versions`)."""

    def __init__(self, pes, layers, f1, nhead: int):
        self.keep = []                 # tensors the raw pointers below refer to

        def mat(w):
            w = f32c(w.detach())
            out = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
            check(lib().lmkd_cast_bf16(ptr(w), ptr(out), w.numel(), stream()), "lmkd_cast_bf16")
            self.keep.append(out)
            return out.data_ptr()

        def vec(v):
            v = f32c(v.detach())
            self.keep.append(v)
            return v.data_ptr()

        enc = _ffi.FusionEncoder()
        enc.nmod, enc.dmod, enc.nhead = len(pes), int(pes[0].LayerNorm.weight.numel()), int(nhead)
        enc.nlayers, enc.dout = len(layers), int(f1.weight.shape[0])
        enc.dff = int(layers[0].linear1.weight.shape[0]) if len(layers) else 8
        enc.ln_eps = float(pes[0].LayerNorm.eps)
        for m, pe in enumerate(pes):
            enc.pe_emb[m] = vec(pe.position_embeddings.weight)
            enc.pe_g[m] = vec(pe.LayerNorm.weight)
            enc.pe_b[m] = vec(pe.LayerNorm.bias)
        self.layers = (_ffi.FusionLayer * max(len(layers), 1))()
        for i, ly in enumerate(layers):
            t = self.layers[i]
            t.w_qkv, t.b_qkv = mat(ly.self_attn.in_proj_weight), vec(ly.self_attn.in_proj_bias)
            t.w_o, t.b_o = mat(ly.self_attn.out_proj.weight), vec(ly.self_attn.out_proj.bias)
            t.w_ff1, t.b_ff1 = mat(ly.linear1.weight), vec(ly.linear1.bias)
            t.w_ff2, t.b_ff2 = mat(ly.linear2.weight), vec(ly.linear2.bias)
            t.ln1_g, t.ln1_b = vec(ly.norm1.weight), vec(ly.norm1.bias)
            t.ln2_g, t.ln2_b = vec(ly.norm2.weight), vec(ly.norm2.bias)
        enc.layers = C.cast(self.layers, C.POINTER(_ffi.FusionLayer))
        enc.w_out, enc.b_out = mat(f1.weight), vec(f1.bias)
        self.enc = enc
        self.max_positions = int(pes[0].position_embeddings.weight.shape[0])


def fusion_encoder_forward(pack: PackedFusionEncoder, xs, shifts=None, out=None, accumulate: bool = False):
    """xs: one [N, L, dmod] CUDA tensor per modality -> [N, L, dout]; `shifts[m]` rolls modality m along the frame
    axis (x'[l] = x[(l + shift) % L]); `accumulate` adds into `out` (the three streams of
    ThreeTRXShiftLoopTime.extract_feature, teacher/code/model.py:1648-1664, are summed that way)."""
    _ffi.poll_status(xs[0].device)
    xs = [f32c(x) for x in xs]
    N, L, dmod = xs[0].shape
    enc = pack.enc
    if len(xs) != enc.nmod or any(tuple(x.shape) != (N, L, dmod) for x in xs) or dmod != enc.dmod:
        raise RuntimeError(f"fusion: expected {enc.nmod} inputs of shape [N, L, {enc.dmod}], got {[tuple(x.shape) for x in xs]}")
    if L > pack.max_positions:
        raise RuntimeError(f"fusion: {L} frames but only {pack.max_positions} position embeddings")
    dev = xs[0].device
    nbytes = lib().lmkd_fusion_workspace_bytes(C.byref(enc), N, L)
    if nbytes == 0:
        raise RuntimeError("lmkd_fusion_workspace_bytes: " + lib().lmkd_last_error().decode())
    ws = _bytes(nbytes, dev)
    if out is None:
        if accumulate:
            raise RuntimeError("fusion: accumulate needs an output tensor")
        out = torch.empty(N, L, enc.dout, dtype=torch.float32, device=dev)
    xp = (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    sh = (C.c_int * len(xs))(*[int(v) for v in (shifts or [0] * len(xs))])
    check(lib().lmkd_fusion_fwd(C.byref(enc), xp, sh, N, L, ptr(out), int(accumulate), ptr(ws), stream()),
          "lmkd_fusion_fwd")
    return out


def upcast_into(src_bf16: torch.Tensor, dst_f32: torch.Tensor) -> torch.Tensor:
    """dst (fp32, preallocated) = src (bf16): widens features that were staged from the host in bf16."""
    if src_bf16.dtype != torch.bfloat16 or dst_f32.dtype != torch.float32 or src_bf16.numel() != dst_f32.numel():
        raise RuntimeError("upcast_into: need a bf16 source and an fp32 destination of the same size")
    check(lib().lmkd_upcast_bf16(ptr(src_bf16), ptr(dst_f32), src_bf16.numel(), stream()), "lmkd_upcast_bf16")
    return dst_f32


def accuracy_count(logits, labels) -> torch.Tensor:
    """#rows with argmax(logits) == label as a device int32 tensor (aggregate_accuracy, utils.py:116-121)."""
    lg = f32c(logits).reshape(-1, logits.shape[-1])
    lab = labels.to(device=lg.device, dtype=torch.int64).reshape(-1).contiguous()
    out = torch.zeros(1, dtype=torch.int32, device=lg.device)
    check(lib().lmkd_accuracy_count(ptr(lg), ptr(lab), lg.shape[0], lg.shape[1], ptr(out), stream()), "lmkd_accuracy")
    return out


# --------------------------------------------------------------------------------------------
# raw GEMM (tests / roofline bench)
# --------------------------------------------------------------------------------------------
def gemm_bf16(A, B, *, a_mn=False, b_mn=False, alpha=1.0, out=None, accumulate=False, block_n=0):
    """C[b] = alpha * A[b] @ B[b]^T with A [b, M, K] (or [b, K, M] if a_mn), B [b, N, K] (or [b, K, N])."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.dim() == 3 and B.dim() == 3
    nb = A.shape[0]
    M, K = (A.shape[2], A.shape[1]) if a_mn else (A.shape[1], A.shape[2])
    N = B.shape[2] if b_mn else B.shape[1]
    if out is None:
        out = torch.empty(nb, M, N, dtype=torch.float32, device=A.device)
    assert A.is_cuda and B.is_cuda and A.stride(2) == 1 and B.stride(2) == 1 and out.stride(2) == 1
    raw = lambda t: C.c_void_p(t.data_ptr())     # strided views are fine: pitches are passed explicitly
    check(lib().lmkd_gemm_bf16(M, N, K, nb, raw(A), int(a_mn), A.stride(1), A.stride(0), raw(B), int(b_mn), B.stride(1),
                               B.stride(0), raw(out), out.stride(1), out.stride(0), alpha, int(accumulate), block_n,
                               stream()), "lmkd_gemm_bf16")
    return out
