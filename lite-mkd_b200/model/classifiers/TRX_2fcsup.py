"""Shipped student / teacher heads (reference: model/classifiers/TRX_2fcsup.py:191-256)."""
import torch
import torch.nn as nn

from .cross_transformer import SupportDK, TemporalCrossTransformer  # noqa: F401
from .TRX_2fc import run_two_heads


class TRX_2fcsup(nn.Module):
    """{'kl': TRX(fc1 features), 'ce': TRX(fc2 features), 'sup': SupportDK(fc2 features)}."""

    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)
        self.supportKD = SupportDK(args)

    def forward(self, context_feature, context_labels, target_feature):
        kl, ce, sup_feat = run_two_heads(self.transformers, context_feature, context_labels, target_feature)
        sup = self.supportKD(sup_feat, context_labels, None)["logits"]
        return {"logits": {"kl": kl, "ce": ce, "sup": sup}}


class TRX_2fcsup_fixed(nn.Module):
    """Frozen teacher: {'kl', 'sup'} on the precomputed multi-modal features."""

    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)
        self.supportKD = SupportDK(args)

    def forward(self, context_feature, context_labels, target_feature):
        with torch.no_grad():
            kl = self.transformers(context_feature, context_labels, target_feature)["logits"]
            sup = self.supportKD(context_feature, context_labels, target_feature)["logits"]
        return {"logits": {"kl": kl, "sup": sup}}
