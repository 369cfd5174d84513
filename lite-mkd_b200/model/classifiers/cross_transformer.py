"""TemporalCrossTransformer on the lmkd CUDA path.

Mirrors the module layout of the reference so checkpoints interchange:
  model/classifiers/TRX.py:24-49   PositionalEncoding  (buffer `pe.pe`, dropout)
  model/classifiers/TRX.py:51-164  TemporalCrossTransformer (k_linear, v_linear, norm_k, norm_v)
  teacher/code/model.py:226-361    the generic-D / generic-cardinality variant
The forward never builds the tuple tensor or loops over classes in Python: it is one call into
`lmkd_trx_fwd` (factored projection GEMM -> tuple assembly + LayerNorm -> class-grouped attention).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from lmkd import ops


class PositionalEncoding(nn.Module):
    """Same buffer as TRX.py:32-41 (0.1-scaled sin/cos table, shape [1, max_len, d_model])."""

    def __init__(self, d_model, dropout, max_len=5000, pe_scale_factor=0.1):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.pe_scale_factor = pe_scale_factor
        position = torch.arange(0, max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(position * div_term) * pe_scale_factor
        table[:, 1::2] = torch.cos(position * div_term) * pe_scale_factor
        self.register_buffer("pe", table.unsqueeze(0))

    def forward(self, x):
        # kept for API parity; the head fuses this add (+ dropout) into its bf16 cast kernel
        return self.dropout(x + self.pe[:, : x.size(1)])


class TemporalCrossTransformer(nn.Module):
    def __init__(self, args, temporal_set_size=2, in_dim=None):
        super().__init__()
        self.args = args
        self.temporal_set_size = temporal_set_size
        d_in = int(in_dim if in_dim is not None else getattr(args, "trans_linear_in_dim", 2048))
        d_out = int(args.trans_linear_out_dim)
        self.in_dim = d_in
        max_len = int(args.seq_len * 1.5)
        self.pe = PositionalEncoding(d_in, float(args.trans_dropout), max_len=max_len)
        self.k_linear = nn.Linear(d_in * temporal_set_size, d_out)
        self.v_linear = nn.Linear(d_in * temporal_set_size, d_out)
        self.norm_k = nn.LayerNorm(d_out)
        self.norm_v = nn.LayerNorm(d_out)          # present in the reference, never applied (TRX.py:110)
        self.class_softmax = torch.nn.Softmax(dim=1)
        tuples, inv_off, inv_idx = ops.tuple_tables(int(args.seq_len), temporal_set_size)
        self.tuples_len = tuples.shape[0]
        # index tables travel with .to(device) but stay out of the state_dict (key contract)
        self.register_buffer("_tuples", tuples, persistent=False)
        self.register_buffer("_inv_off", inv_off, persistent=False)
        self.register_buffer("_inv_idx", inv_idx, persistent=False)
        # device-side dropout counter: bumped (on the device) at every training forward, so a captured
        # CUDA graph draws a fresh PE-dropout mask on every replay
        self.register_buffer("_drop_counter", torch.zeros(1, dtype=torch.int64), persistent=False)

    def head_arguments(self, support_set):
        """Keyword arguments of ops.trx_logits for one call on `support_set` [B,Ns,L,D] (draws the dropout seed)."""
        L = support_set.shape[2]
        if L != int(self.args.seq_len):
            raise RuntimeError(f"seq_len mismatch: features have {L} frames, args.seq_len = {self.args.seq_len}")
        p = float(self.pe.dropout.p) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0.0 else 0
        seed_dev = None
        if p > 0.0 and self._drop_counter.is_cuda:
            self._drop_counter.add_(1)
            seed_dev = self._drop_counter
        return dict(pe=self.pe.pe[0, :L], Wk=self.k_linear.weight, bk=self.k_linear.bias, Wv=self.v_linear.weight,
                    bv=self.v_linear.bias, gamma=self.norm_k.weight, beta=self.norm_k.bias,
                    tables=(self._tuples, self._inv_off, self._inv_idx), card=self.temporal_set_size,
                    way=int(self.args.way), shot=int(self.args.shot), dropout_p=p, seed=seed, ln_eps=self.norm_k.eps,
                    seed_dev=seed_dev)

    def forward_batched(self, support_set, support_labels, queries, with_proto_sim=False):
        """[B,Ns,L,D], [B,Ns], [B,Nq,L,D] -> logits [B,Nq,way] (on the inputs' device); with
        `with_proto_sim` also the [B,Nq,way,way] cosine matrix between per-class query prototypes."""
        h = self.head_arguments(support_set)
        return ops.trx_logits(support_set, support_labels, queries, h.pop("pe"), h.pop("Wk"), h.pop("bk"), h.pop("Wv"),
                              h.pop("bv"), h.pop("gamma"), h.pop("beta"), h.pop("tables"),
                              with_proto_sim=with_proto_sim, **h)

    def forward(self, support_set, support_labels, queries):
        if support_set.dim() == 4:
            return {"logits": self.forward_batched(support_set, support_labels, queries)}
        logits = self.forward_batched(support_set.unsqueeze(0), support_labels.reshape(1, -1), queries.unsqueeze(0))
        return {"logits": logits[0]}


class SupportDK(nn.Module):
    """Support-level inter-prototype logits (TRX_2fcsup.py:162-189); ignores labels like the reference."""

    def __init__(self, args):
        super().__init__()
        self.args = args

    def forward(self, support_set, support_labels, queries):
        way, shot = int(self.args.way), int(self.args.shot)
        if support_set.dim() == 4:
            return {"logits": ops.support_dk(support_set, way, shot)}
        return {"logits": ops.support_dk(support_set.unsqueeze(0), way, shot)[0]}
