"""STRM heads on the lmkd CUDA path (reference: model/classifiers/strm_res18_sup.py:162-325,
strmclassifiers.py:162-289, strmclassifiers_res18.py:162-288 -- the three files carry the same DistanceLoss).

DistanceLoss = dropout -> all ordered frame tuples -> clsW: Linear(c*D -> D/2) + ReLU on every tuple -> per class the
Euclidean distance of every query tuple to its NEAREST support tuple of that class -> mean over the query's tuples,
negated.  Here it is one call into `lmkd_strm_dist_fwd`: factored tuple MLP (one tcgen05 GEMM on frames), tuple
assembly + ReLU, and a tuple-to-tuple distance GEMM whose epilogue keeps only the per-row arg-min.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from lmkd import ops

from .cross_transformer import SupportDK, TemporalCrossTransformer


class DistanceLoss(nn.Module):
    """Query-class similarity on the patch-enriched features (strm_res18_sup.py:162-243)."""

    def __init__(self, args, temporal_set_size=2):
        super().__init__()
        self.args = args
        self.temporal_set_size = temporal_set_size
        d_in = int(getattr(args, "trans_linear_in_dim", 2048))
        self.dropout = nn.Dropout(p=0.1)                                    # :170, fixed at 0.1 in the reference
        self.clsW = nn.Linear(d_in * temporal_set_size, d_in // 2)         # :179
        self.relu = torch.nn.ReLU()
        tuples, inv_off, inv_idx = ops.tuple_tables(int(args.seq_len), temporal_set_size)
        self.tuples_len = tuples.shape[0]
        self.register_buffer("_tuples", tuples, persistent=False)
        self.register_buffer("_inv_off", inv_off, persistent=False)
        self.register_buffer("_inv_idx", inv_idx, persistent=False)
        self.register_buffer("_drop_counter", torch.zeros(1, dtype=torch.int64), persistent=False)

    def forward_batched(self, support_set, support_labels, queries):
        p = float(self.dropout.p) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0.0 else 0
        seed_dev = None
        if p > 0.0 and self._drop_counter.is_cuda:
            self._drop_counter.add_(1)
            seed_dev = self._drop_counter
        return ops.strm_distance_logits(support_set, support_labels, queries, self.clsW.weight, self.clsW.bias,
                                        (self._tuples, self._inv_off, self._inv_idx), card=self.temporal_set_size,
                                        way=int(self.args.way), shot=int(self.args.shot), dropout_p=p, seed=seed,
                                        seed_dev=seed_dev)

    def forward(self, support_set, support_labels, queries, device=None):
        # `device` is accepted for call compatibility (:184); the logits stay where the features are
        if support_set.dim() == 4:
            return {"logits": self.forward_batched(support_set, support_labels, queries)}
        lg = self.forward_batched(support_set.unsqueeze(0), support_labels.reshape(1, -1), queries.unsqueeze(0))
        return {"logits": lg[0]}


class _StrmTwoStream(nn.Module):
    """strmclassifiers / strmclassifiers_resnet18: DistanceLoss on the 'distance' features, TRX on the 'trx' ones."""

    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)
        self.DistanceLoss = DistanceLoss(args, 2)

    def forward(self, context_feature, context_labels, target_feature):
        dev = getattr(self.args, "device", None)
        pat = self.DistanceLoss(context_feature["distance"], context_labels, target_feature["distance"], dev)["logits"]
        fr = self.transformers(context_feature["trx"], context_labels, target_feature["trx"])["logits"]
        return {"logits": {"pat": pat, "fr": fr}}


class strmclassifiers(_StrmTwoStream):                   # strmclassifiers.py:257-289
    pass


class strmclassifiers_resnet18(_StrmTwoStream):          # strmclassifiers_res18.py:257-288
    pass


class strmclassifiers_resnet18_sup(nn.Module):           # strm_res18_sup.py:288-325
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)
        self.DistanceLoss = DistanceLoss(args, 2)
        self.SupportDK = SupportDK(args)

    def forward(self, context_feature, context_labels, target_feature):
        dev = getattr(self.args, "device", None)
        pat = self.DistanceLoss(context_feature["distance"], context_labels, target_feature["distance"], dev)["logits"]
        fr1 = self.transformers(context_feature["trx1"], context_labels, target_feature["trx1"])["logits"]
        fr2 = self.transformers(context_feature["trx2"], context_labels, target_feature["trx2"])["logits"]
        sup = self.SupportDK(context_feature["trx2"], context_labels, target_feature["trx2"])["logits"]
        return {"logits": {"pat": pat, "fr1": fr1, "sup": sup, "fr2": fr2}}
