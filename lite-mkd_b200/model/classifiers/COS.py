"""CosDistance (reference: model/classifiers/COS.py:24-62) — despite its name it is the same
frame-mean Euclidean computation as e_dist, and it returns the logits tensor itself, not a dict."""
import torch.nn as nn

from .e_dist import edist_forward


class CosDistance(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args

    def forward(self, support_set, support_labels, queries):
        return edist_forward(self.args, support_set, support_labels, queries)
