"""TRX heads (reference: model/classifiers/TRX.py:167-211)."""
import torch
import torch.nn as nn

from lmkd import ops

from .cross_transformer import PositionalEncoding, TemporalCrossTransformer  # noqa: F401


class TRX(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)

    def forward(self, context_feature, context_labels, target_feature):
        return self.transformers(context_feature, context_labels, target_feature)


class TRX_fixed(nn.Module):
    """Frozen teacher wrapper; also accepts flat features like the reference (:205-206)."""

    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = TemporalCrossTransformer(args, 2)

    def forward(self, context_feature, context_labels, target_feature):
        with torch.no_grad():
            if context_feature.dim() != 4:
                context_feature = context_feature.reshape(-1, self.args.seq_len, self.transformers.in_dim)
                target_feature = target_feature.reshape(-1, self.args.seq_len, self.transformers.in_dim)
            logits = self.transformers(context_feature, context_labels, target_feature)["logits"]
        return {"logits": logits}


class TrxBranch(nn.Module):
    """Multi-cardinality TRX (teacher/code/model.py:1094-1128): one transformer per entry of
    args.temp_set, logits averaged.  Classifier-style argument order (support, labels, query);
    returns [Nq, way] (the reference adds a leading sample dim of 1)."""

    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.transformers = nn.ModuleList([TemporalCrossTransformer(args, s) for s in args.temp_set])

    def forward(self, context_feature, context_labels, target_feature):
        # one autograd node for the whole branch: the cardinalities' feature gradients are summed inside the backward
        # kernels and the mean over cardinalities never leaves the op (lmkd.ops._TrxBranchFn)
        single = context_feature.dim() != 4
        if single:
            context_feature, context_labels, target_feature = (context_feature.unsqueeze(0), context_labels.reshape(1, -1),
                                                               target_feature.unsqueeze(0))
        heads = [t.head_arguments(context_feature) for t in self.transformers]
        logits = ops.trx_branch_logits(context_feature, context_labels, target_feature, heads)
        return {"logits": logits[0] if single else logits}
