"""Two-feature-head TRX wrapper (reference: model/classifiers/TRX_2fc.py:163-192)."""
import torch
import torch.nn as nn

from .cross_transformer import TemporalCrossTransformer


def run_two_heads(transformer, context_feature, context_labels, target_feature):
    """Both feature heads share the transformer weights, so they run as one batch of episodes."""
    c1, c2 = context_feature["context_features_1"], context_feature["context_features_2"]
    t1, t2 = target_feature["target_features_1"], target_feature["target_features_2"]
    if c1.dim() == 3:
        sup = torch.stack([c1, c2])
        qry = torch.stack([t1, t2])
        lab = torch.stack([context_labels, context_labels])
        out = transformer.forward_batched(sup, lab, qry)
        return out[0], out[1], c2
    B = c1.shape[0]
    out = transformer.forward_batched(torch.cat([c1, c2]), torch.cat([context_labels, context_labels]),
                                      torch.cat([t1, t2]))
    return out[:B], out[B:], c2


class TRX_2fc(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)

    def forward(self, context_feature, context_labels, target_feature):
        l1, l2, _ = run_two_heads(self.transformers, context_feature, context_labels, target_feature)
        return {"logits": {"fc_1": l1, "fc_2": l2}}
