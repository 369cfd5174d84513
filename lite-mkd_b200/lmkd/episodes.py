"""Seeded synthetic episodes shaped like the BASELINE.json configs.

Class-structured features: per episode one centroid per class, every video = its class centroid
+ 0.5 * N(0,1); teacher features = the same centroids + independent noise, optionally the sum of
three modality streams (mirrors extract_feature, teacher/code/model.py:1648-1664).  Labels are
shuffled like video_reader.py:454-460 and stored float for the classifier, int64 for the loss.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

SEED = 3483  # the reference's own seeding constant (model/classifiers/TRX.py:18-21)


@dataclass
class EpisodeBatch:
    support: torch.Tensor          # [B, Ns, L, D] student features
    query: torch.Tensor            # [B, Nq, L, D]
    support_labels: torch.Tensor   # [B, Ns] float
    query_labels: torch.Tensor     # [B, Nq] int64
    teacher_support: torch.Tensor  # [B, Ns, L, Dt]
    teacher_query: torch.Tensor    # [B, Nq, L, Dt]

    def to(self, device, non_blocking=False):
        return EpisodeBatch(*[t.to(device, non_blocking=non_blocking) for t in self.tensors()])

    def pin(self):
        return EpisodeBatch(*[t.pin_memory() for t in self.tensors()])

    def tensors(self):
        return (self.support, self.query, self.support_labels, self.query_labels, self.teacher_support,
                self.teacher_query)

    def slice(self, lo, hi):
        return EpisodeBatch(*[t[lo:hi] for t in self.tensors()])


def make_episodes(B, way=5, shot=5, query_per_class=5, L=8, D=2048, teacher_dim=2048, *, noise=0.5,
                  modalities=1, shuffle=True, class_sorted_support=False, seed=SEED, device="cpu",
                  dtype=torch.float32, separation=1.0) -> EpisodeBatch:
    """`separation` scales the class centroids (std of their entries).  1.0 gives episodes every head classifies
    with margins of hundreds of logit units (argmax tests); with random-initialised heads the softmax of such
    logits is exactly one-hot in fp32 and every loss gradient w.r.t. the head parameters underflows to 0 -- the
    benchmark uses 0.1 so that the backward pass and the optimizer work on live gradients."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    dev = torch.device(device)

    def randn(*shape):
        # generated on the target device when possible (big tensors), seeded either way
        if dev.type == "cuda":
            gg = torch.Generator(device=dev).manual_seed(int(torch.randint(0, 2 ** 62, (1,), generator=g).item()))
            return torch.randn(*shape, generator=gg, device=dev, dtype=dtype)
        return torch.randn(*shape, generator=g, dtype=dtype)

    Ns, Nq = way * shot, way * query_per_class
    s_lab = torch.arange(way).repeat_interleave(shot).repeat(B, 1)
    q_lab = torch.arange(way).repeat_interleave(query_per_class).repeat(B, 1)
    if shuffle:
        if not class_sorted_support:
            s_lab = torch.stack([r[torch.randperm(Ns, generator=g)] for r in s_lab])
        q_lab = torch.stack([r[torch.randperm(Nq, generator=g)] for r in q_lab])
    s_lab, q_lab = s_lab.to(dev), q_lab.to(dev)
    cent = randn(B, way, L, D) * separation
    idx_s = s_lab[:, :, None, None].expand(B, Ns, L, D)
    idx_q = q_lab[:, :, None, None].expand(B, Nq, L, D)
    support = torch.gather(cent, 1, idx_s) + noise * randn(B, Ns, L, D)
    query = torch.gather(cent, 1, idx_q) + noise * randn(B, Nq, L, D)
    if teacher_dim == D:
        tcent = cent
    else:
        tcent = randn(B, way, L, teacher_dim) * separation
    tidx_s = s_lab[:, :, None, None].expand(B, Ns, L, teacher_dim)
    tidx_q = q_lab[:, :, None, None].expand(B, Nq, L, teacher_dim)
    t_support = torch.gather(tcent, 1, tidx_s)
    t_query = torch.gather(tcent, 1, tidx_q)
    ts, tq = torch.zeros_like(t_support), torch.zeros_like(t_query)
    for _ in range(modalities):   # sum of modality streams, each centroid/m + noise/sqrt(m)
        ts = ts + t_support / modalities + (noise / modalities ** 0.5) * randn(B, Ns, L, teacher_dim)
        tq = tq + t_query / modalities + (noise / modalities ** 0.5) * randn(B, Nq, L, teacher_dim)
    return EpisodeBatch(support, query, s_lab.float(), q_lab.long(), ts, tq)
