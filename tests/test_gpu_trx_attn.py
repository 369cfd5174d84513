"""Fused TRX attention kernel (lmkd_trx_attn_fwd) against a plain torch fp32 restatement of
model/classifiers/TRX.py:120-148 on the same bf16 operands (scores, per-class softmax, prototype, distance).

Tolerances: the kernel keeps exp(score - max) in bf16 (8 mantissa bits) as the A operand of P.V and accumulates in
fp32, so prototypes agree to ~2^-9 relative; the row sums of squares to 1e-2 relative (north_star: logits 1e-2).
"""
import ctypes as C
import math

import pytest
import torch


from conftest import allclose, record_error

pytestmark = pytest.mark.gpu


def _shape(B, way, shot, Nq, L, card, d):
    from lmkd._ffi import TrxShape
    return TrxShape(B, way * shot, Nq, L, 64, d, card, way, shot, 0.0, 0, None, 1e-5)


def _reference(kq, vq, ks, vs, cnt, T, d):
    """fp32 torch: returns diff [B,way,NqT,d], rowred, rowdot, linv [B,way,NqT], ptilde [B,NqT,way,KTp]."""
    B, NqT, _ = kq.shape
    way, KTp = ks.shape[1], ks.shape[2]
    kq, vq, ks, vs = kq.float(), vq.float(), ks.float(), vs.float()
    S = torch.einsum("bmd,bcnd->bcmn", kq, ks) / math.sqrt(d)
    valid = (torch.arange(KTp, device=kq.device)[None, None, :] < (cnt.view(B, way, 1) * T)).view(B, way, 1, KTp)
    S = S.masked_fill(~valid, float("-inf"))
    mx = S.amax(-1, keepdim=True)
    mx = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
    E = torch.exp(S - mx)
    E = torch.where(valid.expand_as(E), E, torch.zeros_like(E))
    l = E.sum(-1, keepdim=True)
    linv = torch.where(l > 0, 1.0 / l, torch.zeros_like(l))
    O = torch.einsum("bcmn,bcnd->bcmd", E * linv, vs)
    diff = vq[:, None] - O
    return diff, (diff * diff).sum(-1), (diff * O).sum(-1), linv.squeeze(-1), E.permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,way,shot,Nq,L,card,d,ragged", [
    (2, 5, 5, 25, 8, 2, 1152, False),     # cfg2 episode, pairs: KTp 144 (one MMA along N)
    (2, 5, 5, 25, 8, 3, 1152, False),     # cfg2 episode, triples: KTp 288 (two MMAs along N), 1400 rows = 10.94 tiles
    (3, 5, 1, 7, 8, 2, 128, False),       # cfg1-like: KTp 32
    (2, 4, 3, 6, 8, 2, 256, True),        # classes with fewer supports than `shot` and an empty class
    (1, 3, 6, 5, 8, 3, 64, False),        # KTp 336: two unequal N halves (176 + 160)
])
def test_fused_attention_matches_torch(B, way, shot, Nq, L, card, d, ragged):
    from lmkd._ffi import check, lib, ptr, stream
    dev = torch.device("cuda:0")
    T = math.comb(L, card)
    KT = shot * T
    KTp = (KT + 15) // 16 * 16
    NqT = Nq * T
    sh = _shape(B, way, shot, Nq, L, card, d)
    assert lib().lmkd_trx_attn_fused_fits(C.byref(sh)) == 1
    g = torch.Generator(device="cpu").manual_seed(100 + card + d)
    # LayerNorm-like keys (unit variance), values of similar scale, scores spread over a few units
    kq = torch.randn(B, NqT, d, generator=g).to(dev).bfloat16()
    vq = torch.randn(B, NqT, d, generator=g).to(dev).bfloat16()
    ks = torch.randn(B, way, KTp, d, generator=g).to(dev).bfloat16()
    vs = torch.randn(B, way, KTp, d, generator=g).to(dev).bfloat16()
    cnt = torch.full((B, way), shot, dtype=torch.int32)
    if ragged:
        cnt[0, 1] = shot - 1
        cnt[1, 0] = 0
        cnt[1, 2] = 1
    cnt = cnt.to(dev)
    rows = torch.arange(KTp, device=dev)[None, None, :, None]
    pad = rows >= (cnt.view(B, way, 1, 1) * T)
    ks = ks.masked_fill(pad, 0).contiguous()       # the tuple kernel zeroes these rows (trx_zero_pad_rows)
    vs = vs.masked_fill(pad, 0).contiguous()
    dq = torch.full((B, way, NqT, d), float("nan"), device=dev).bfloat16()
    patt = torch.full((B, NqT, way * KTp), float("nan"), device=dev).bfloat16()
    rowred = torch.zeros(B, way, NqT, device=dev)
    rowdot = torch.zeros(B, way, NqT, device=dev)
    linv = torch.full((B, way, NqT), float("nan"), device=dev)
    check(lib().lmkd_trx_attn_fwd(C.byref(sh), ptr(kq), ptr(vq), ptr(ks), ptr(vs), ptr(cnt), ptr(dq), ptr(patt),
                                  ptr(rowred), ptr(rowdot), ptr(linv), stream()), "lmkd_trx_attn_fwd")
    torch.cuda.synchronize()
    diff_r, rowred_r, rowdot_r, linv_r, pt_r = _reference(kq, vq, ks, vs, cnt, T, d)
    tag = f"trx_attn[c{card},d{d},KTp{KTp}]"
    e_diff = ((dq.float() - diff_r).norm() / diff_r.norm()).item()
    e_red = ((rowred - rowred_r).abs() / rowred_r.abs().clamp_min(1e-6)).max().item()
    e_dot = ((rowdot - rowdot_r).norm() / rowdot_r.norm().clamp_min(1e-6)).item()
    e_linv = ((linv - linv_r).abs() / linv_r.abs().clamp_min(1e-12)).max().item()
    pt = patt.float().view(B, NqT, way, KTp)
    e_p = (pt - pt_r).abs().max().item()
    record_error(tag, diff_rel_l2=e_diff, rowred_max_rel=e_red, rowdot_rel_l2=e_dot, linv_max_rel=e_linv,
                 ptilde_max_abs=e_p)
    assert torch.isfinite(dq.float()).all() and torch.isfinite(patt.float()).all()
    assert e_diff < 6e-3, e_diff            # bf16 output rounding (2^-9) dominates
    assert e_red < 1e-2, e_red
    assert e_dot < 1e-2, e_dot
    assert e_linv < 2e-3, e_linv            # ex2.approx + fp32 summation order
    assert e_p < 8e-3, e_p                  # bf16 rounding of values in [0, 1]
    # the second call without the optional outputs (no-grad pass) gives the same row sums
    rowred2 = torch.zeros_like(rowred)
    check(lib().lmkd_trx_attn_fwd(C.byref(sh), ptr(kq), ptr(vq), ptr(ks), ptr(vs), ptr(cnt), None, None,
                                  ptr(rowred2), None, None, stream()), "lmkd_trx_attn_fwd")
    torch.cuda.synchronize()
    assert allclose(rowred2, rowred, rtol=1e-5, atol=1e-5)


def test_fused_attention_many_items_is_deterministic_per_item():
    """More work items than SMs (every CTA loops, rings wrap with both phases): an episode repeated along the
    batch must give identical results in every copy."""
    from lmkd._ffi import check, lib, ptr, stream
    dev = torch.device("cuda:0")
    B, way, shot, Nq, L, card, d = 12, 5, 5, 25, 8, 3, 1152
    T = math.comb(L, card)
    KTp = (shot * T + 15) // 16 * 16
    NqT = Nq * T
    sh = _shape(B, way, shot, Nq, L, card, d)
    g = torch.Generator(device="cpu").manual_seed(5)
    kq = torch.randn(1, NqT, d, generator=g).to(dev).bfloat16().expand(B, -1, -1).contiguous()
    vq = torch.randn(1, NqT, d, generator=g).to(dev).bfloat16().expand(B, -1, -1).contiguous()
    ks = torch.randn(1, way, KTp, d, generator=g).to(dev).bfloat16()
    vs = torch.randn(1, way, KTp, d, generator=g).to(dev).bfloat16()
    ks[:, :, shot * T:] = 0
    vs[:, :, shot * T:] = 0
    ks = ks.expand(B, -1, -1, -1).contiguous()
    vs = vs.expand(B, -1, -1, -1).contiguous()
    cnt = torch.full((B, way), shot, dtype=torch.int32, device=dev)
    dq = torch.zeros(B, way, NqT, d, device=dev).bfloat16()
    patt = torch.zeros(B, NqT, way * KTp, device=dev).bfloat16()
    rowred = torch.zeros(B, way, NqT, device=dev)
    rowdot = torch.zeros(B, way, NqT, device=dev)
    linv = torch.zeros(B, way, NqT, device=dev)
    check(lib().lmkd_trx_attn_fwd(C.byref(sh), ptr(kq), ptr(vq), ptr(ks), ptr(vs), ptr(cnt), ptr(dq), ptr(patt),
                                  ptr(rowred), ptr(rowdot), ptr(linv), stream()), "lmkd_trx_attn_fwd")
    torch.cuda.synchronize()
    for b in range(1, B):
        assert torch.equal(dq[b], dq[0]) and torch.equal(patt[b], patt[0]) and torch.equal(linv[b], linv[0])
        assert allclose(rowred[b], rowred[0], rtol=1e-6, atol=0)     # two atomic adds per row: order may differ
