"""Oracle restatement of the matching heads (test infrastructure, see oracle/__init__.py).

All functions are pure torch on whatever device/dtype the inputs carry (tests use CPU
fp32/fp64); gradients come from autograd.  Citations are reference paths (file:line).
"""
from __future__ import annotations

import math
from itertools import combinations

import torch

# --------------------------------------------------------------------------------------
# Frame similarity + OTAM  (teacher/code/model.py:3260-3343)
# --------------------------------------------------------------------------------------


def cos_sim(x: torch.Tensor, y: torch.Tensor, eps: float = 0.01) -> torch.Tensor:
    """teacher/code/model.py:3260-3269 — <x,y> / (|x|·|y| + eps); eps sits on the norm product."""
    dots = x @ y.transpose(-1, -2)
    denom = x.norm(dim=-1)[..., :, None] * y.norm(dim=-1)[..., None, :] + eps
    return dots / denom


def _softmin(vals, lbda: float, stable: bool) -> torch.Tensor:
    a = torch.stack(vals, dim=-1)
    if stable:
        return -lbda * torch.logsumexp(-a / lbda, dim=-1)
    # the reference's literal form: -λ·log Σ exp(-a/λ), no max subtraction
    return -lbda * torch.log(torch.exp(-a / lbda).sum(dim=-1))


def otam_cum_dist(d: torch.Tensor, lbda: float = 0.1, stable: bool = False) -> torch.Tensor:
    """teacher/code/model.py:3271-3299 for d[..., L, M] -> [...].

    The table is the distance matrix padded with a zero column on either side.
    Row 0 is a plain running sum (:3280-3283); column 0 is never written and stays 0;
    column 1 and the last (padding) column use a three-way soft-min (:3288, :3296);
    interior columns a two-way soft-min (:3291-3292).
    """
    L, M = d.shape[-2], d.shape[-1]
    zero = torch.zeros_like(d[..., 0, 0])
    prev = [zero]
    for m in range(1, M + 1):
        prev.append(prev[-1] + d[..., 0, m - 1])
    prev.append(prev[-1])                       # padded column contributes 0
    for l in range(1, L):
        cur = [zero]
        cur.append(d[..., l, 0] + _softmin([prev[0], prev[1], cur[0]], lbda, stable))
        for m in range(2, M + 1):
            cur.append(d[..., l, m - 1] + _softmin([prev[m - 1], cur[m - 1]], lbda, stable))
        cur.append(_softmin([prev[M], prev[M + 1], cur[M]], lbda, stable))
        prev = cur
    return prev[M + 1]


def otam_cum_dist_stable(d: torch.Tensor, lbda: float = 0.1) -> torch.Tensor:
    """Same recurrence with a log-sum-exp soft-min: identical wherever the reference is finite,
    and finite where the reference's exp(-c/λ) underflows (L >= 12)."""
    return otam_cum_dist(d, lbda, stable=True)


def otam_pair_dists(support: torch.Tensor, query: torch.Tensor, lbda: float = 0.1,
                    stable: bool = True) -> torch.Tensor:
    """teacher/code/model.py:3329-3338 — [Nq, Ns] bidirectional OTAM distance."""
    nq, L, _ = query.shape
    ns, M, _ = support.shape
    dist = 1.0 - cos_sim(query.reshape(nq * L, -1), support.reshape(ns * M, -1))
    d = dist.reshape(nq, L, ns, M).permute(0, 2, 1, 3)
    return otam_cum_dist(d, lbda, stable) + otam_cum_dist(d.transpose(-1, -2), lbda, stable)


def otam_logits(support: torch.Tensor, labels: torch.Tensor, query: torch.Tensor,
                lbda: float = 0.1, stable: bool = True) -> torch.Tensor:
    """teacher/code/model.py:3319-3343 (CNN_OTAM.forward) -> [Nq, n_classes] probabilities.

    Columns follow the sorted unique labels (:3340); the output is softmax(-class mean dist)
    over classes (:3343, implicit dim=1).  NaN guard (:3322-3324): zeros [Nq, 5].
    """
    if torch.isnan(support).any():
        return torch.zeros(query.shape[0], 5, dtype=query.dtype)
    cum = otam_pair_dists(support, query, lbda, stable)
    classes = torch.unique(labels)
    cls = torch.stack([cum[:, labels == c].mean(dim=1) for c in classes], dim=1)
    return torch.softmax(-cls, dim=1)


# --------------------------------------------------------------------------------------
# TRX  (model/classifiers/TRX.py:24-164; generic D / cardinality: teacher/code/model.py:226-361)
# --------------------------------------------------------------------------------------


def positional_encoding_table(max_len: int, d_model: int, scale: float = 0.1) -> torch.Tensor:
    """model/classifiers/TRX.py:32-41 — computed in fp32 exactly as the reference buffer is."""
    pos = torch.arange(0, max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div) * scale
    pe[:, 1::2] = torch.cos(pos * div) * scale
    return pe


def frame_tuples(seq_len: int, card: int):
    """model/classifiers/TRX.py:70-73 — lexicographic c-combinations of range(L)."""
    return list(combinations(range(seq_len), card))


def _project_tuples(x, tuples, W, b):
    """TRX.py:90-104 — gather c frames per tuple, concatenate, apply Linear(c·D -> d)."""
    idx = torch.tensor(tuples, dtype=torch.long)             # [T, c]
    n = x.shape[0]
    u = x[:, idx, :].reshape(n, idx.shape[0], -1)             # [N, T, c·D]
    return u @ W.t() + b


def _layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def trx_class_prototypes(support, labels, query, Wk, bk, Wv, bv, gk, bek, card,
                         pe=None, ln_eps: float = 1e-5):
    """Returns (classes, v_q [Nq,T,d], protos list of [Nq,T,d]) following TRX.py:84-138."""
    L = support.shape[1]
    if pe is None:
        pe = positional_encoding_table(int(L * 1.5), support.shape[2]).to(support.dtype)
    s = support + pe[:L]
    q = query + pe[:L]
    tuples = frame_tuples(L, card)
    d_out = Wk.shape[0]
    ks = _layer_norm(_project_tuples(s, tuples, Wk, bk), gk, bek, ln_eps)   # TRX.py:107
    kq = _layer_norm(_project_tuples(q, tuples, Wk, bk), gk, bek, ln_eps)   # TRX.py:108
    vs = _project_tuples(s, tuples, Wv, bv)                                  # values un-normed :110
    vq = _project_tuples(q, tuples, Wv, bv)
    classes = torch.unique(labels)
    protos = []
    nq, T = kq.shape[0], kq.shape[1]
    for c in classes:
        sel = labels == c
        ck = ks[sel].reshape(-1, d_out)                       # [K·T, d]
        cv = vs[sel].reshape(-1, d_out)
        scores = (kq.reshape(nq * T, d_out) @ ck.t()) / math.sqrt(d_out)    # TRX.py:125
        p = torch.softmax(scores, dim=-1)                     # over all K·T pairs, TRX.py:127-134
        protos.append((p @ cv).reshape(nq, T, d_out))         # TRX.py:137-138
    return classes, vq, protos


def trx_logits(support, labels, query, Wk, bk, Wv, bv, gk, bek, card, way,
               pe=None, ln_eps: float = 1e-5) -> torch.Tensor:
    """TemporalCrossTransformer.forward in eval mode (TRX.py:75-152) -> [Nq, way]."""
    classes, vq, protos = trx_class_prototypes(support, labels, query, Wk, bk, Wv, bv, gk, bek,
                                               card, pe, ln_eps)
    T = vq.shape[1]
    cols = [None] * way
    for c, proto in zip(classes, protos):
        diff = vq - proto
        cols[int(c.item())] = -(diff * diff).sum(dim=(1, 2)) / T          # TRX.py:141-148
    zero = torch.zeros(vq.shape[0], dtype=vq.dtype)
    return torch.stack([z if z is not None else zero for z in cols], dim=1)


def trx_branch_logits(support, labels, query, heads, way, pe=None) -> torch.Tensor:
    """teacher/code/model.py:1109-1127 (TrxBranch.forward): mean over cardinalities.

    `heads` is a list of dicts with keys Wk,bk,Wv,bv,gk,bek,card.  Returned as [Nq, way]
    (the reference adds a leading sample dim of 1, :1125).
    """
    outs = [trx_logits(support, labels, query, h["Wk"], h["bk"], h["Wv"], h["bv"], h["gk"],
                       h["bek"], h["card"], way, pe) for h in heads]
    return torch.stack(outs, dim=-1).mean(dim=-1)


def trx_sup_outputs(support, labels, query, Wk, bk, Wv, bv, gk, bek, card, way, pe=None):
    """model/classifiers/TRX_sup.py:114-179 -> (support_sim [Nq,way,way], query logits [Nq,way])."""
    classes, vq, protos = trx_class_prototypes(support, labels, query, Wk, bk, Wv, bv, gk, bek,
                                               card, pe)
    nq, T, d = vq.shape
    stacked = torch.zeros(nq, T * d, way, dtype=vq.dtype)
    logits = torch.zeros(nq, way, dtype=vq.dtype)
    cols, lcols = [None] * way, [None] * way
    for c, proto in zip(classes, protos):
        cols[int(c.item())] = proto.reshape(nq, T * d)
        diff = vq - proto
        lcols[int(c.item())] = -(diff * diff).sum(dim=(1, 2)) / T
    zp = torch.zeros(nq, T * d, dtype=vq.dtype)
    stacked = torch.stack([c if c is not None else zp for c in cols], dim=2)
    logits = torch.stack([c if c is not None else torch.zeros(nq, dtype=vq.dtype) for c in lcols], 1)
    sim = torch.nn.functional.cosine_similarity(stacked.unsqueeze(2), stacked.unsqueeze(3), dim=1)
    return sim, logits


def support_dk(support: torch.Tensor, way: int, shot: int, seq_len: int) -> torch.Tensor:
    """model/classifiers/TRX_2fcsup.py:175-189 — ignores labels, assumes class-sorted supports.

    out[i, m] = -|proto_i - proto_n|_F^2 / L for the m-th n != i.  The reference hard-codes
    5x4; this restatement uses way x (way-1), identical at way=5.
    """
    proto = support.reshape(way, shot, seq_len, -1).mean(dim=1)
    rows = []
    for i in range(way):
        rows.append(torch.stack([-((proto[i] - proto[n]) ** 2).sum() / seq_len
                                 for n in range(way) if n != i]))
    return torch.stack(rows)


def e_dist_logits(support: torch.Tensor, labels: torch.Tensor, query: torch.Tensor, way: int) -> torch.Tensor:
    """model/classifiers/e_dist.py:22-61 (and COS.py:29-62, which is the same Euclidean computation):
    frame-mean embeddings, cdist(query, supports of class c), mean over the class -> negated."""
    qm = query.mean(dim=1)
    sm = support.mean(dim=1)
    cols = [None] * way
    for c in torch.unique(labels):
        dm = torch.cdist(qm, sm[labels == c], p=2)
        cols[int(c.item())] = -dm.mean(dim=1)
    zero = torch.zeros(qm.shape[0], dtype=qm.dtype)
    return torch.stack([z if z is not None else zero for z in cols], dim=1)


def strm_distance_logits(support: torch.Tensor, labels: torch.Tensor, query: torch.Tensor, W: torch.Tensor,
                         b: torch.Tensor, card: int, way: int) -> torch.Tensor:
    """model/classifiers/strm_res18_sup.py:184-243 (DistanceLoss.forward, eval mode: dropout is the identity):
    tuples of `card` frames concatenated, clsW + ReLU on every tuple (:204-207, :221-224), per class torch.cdist of
    the query tuples against that class's support tuples (:227), minimum over the support tuples (:230), mean over
    the query's tuples, negated (:233-237).  Classes without supports keep 0 (:212)."""
    tuples = frame_tuples(support.shape[1], card)
    nq, T = query.shape[0], len(tuples)
    es = torch.relu(_project_tuples(support, tuples, W, b))          # [Ns, T, dm]
    eq = torch.relu(_project_tuples(query, tuples, W, b)).reshape(nq * T, -1)
    cols = [None] * way
    for c in torch.unique(labels):
        ck = es[labels == c].reshape(-1, es.shape[-1])
        dm = torch.cdist(eq, ck)                                      # [Nq*T, K*T]
        cols[int(c.item())] = -dm.min(dim=1)[0].reshape(nq, T).mean(dim=1)
    zero = torch.zeros(nq, dtype=query.dtype)
    return torch.stack([z if z is not None else zero for z in cols], dim=1)
