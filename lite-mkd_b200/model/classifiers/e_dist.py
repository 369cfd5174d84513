"""Frame-mean Euclidean head (reference: model/classifiers/e_dist.py:16-61).  The reference reshapes
with literals (-1, 8, 2048); here the frame count and width come from args / the tensor."""
import torch.nn as nn

from lmkd import ops


def edist_forward(args, support_set, support_labels, queries):
    way = int(args.way)
    if support_set.dim() == 4:
        return ops.edist_logits(support_set, support_labels, queries, way)
    L = int(args.seq_len)
    if support_set.dim() == 2:                       # flat [N*L, D] features, as the reference accepts
        support_set = support_set.reshape(-1, L, support_set.shape[-1])
        queries = queries.reshape(-1, L, queries.shape[-1])
    return ops.edist_logits(support_set.unsqueeze(0), support_labels.reshape(1, -1), queries.unsqueeze(0), way)[0]


class e_dist(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args

    def forward(self, support_set, support_labels, queries):
        return {"logits": edist_forward(self.args, support_set, support_labels, queries)}
