"""TRX variant that also returns the cosine similarity between the per-class query prototypes
(reference: model/classifiers/TRX_sup.py:74-229; pairs with Distiller.support_sim)."""
import torch
import torch.nn as nn

from .cross_transformer import TemporalCrossTransformer


def _run(transformer, context_feature, context_labels, target_feature):
    if context_feature.dim() == 4:
        logits, sim = transformer.forward_batched(context_feature, context_labels, target_feature, with_proto_sim=True)
        return {"logits": {"support_set": sim, "query": logits}}
    logits, sim = transformer.forward_batched(context_feature.unsqueeze(0), context_labels.reshape(1, -1),
                                              target_feature.unsqueeze(0), with_proto_sim=True)
    return {"logits": {"support_set": sim[0], "query": logits[0]}}


class TRX_sup(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)

    def forward(self, context_feature, context_labels, target_feature):
        return _run(self.transformers, context_feature, context_labels, target_feature)


class TRX_sup_fixed(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)

    def forward(self, context_feature, context_labels, target_feature):
        with torch.no_grad():
            return _run(self.transformers, context_feature, context_labels, target_feature)
