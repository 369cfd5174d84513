"""Two-head / support-level wrappers of the Euclidean head
(reference: model/classifiers/e_dist_fc2.py:106-231)."""
import torch.nn as nn

from .cross_transformer import SupportDK
from .e_dist import e_dist


class e_dist_fc2(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.e_dict = e_dist(args)

    def forward(self, context_feature, context_labels, target_feature):
        l1 = self.e_dict(context_feature["context_features_1"], context_labels, target_feature["target_features_1"])
        l2 = self.e_dict(context_feature["context_features_2"], context_labels, target_feature["target_features_2"])
        return {"logits": {"fc_1": l1["logits"], "fc_2": l2["logits"]}}


class e_dist_fc2_sup(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.e_dict = e_dist(args)
        self.supportKD = SupportDK(args)

    def forward(self, context_feature, context_labels, target_feature):
        c1, c2 = context_feature["context_features_1"], context_feature["context_features_2"]
        t1, t2 = target_feature["target_features_1"], target_feature["target_features_2"]
        return {"logits": {"kl": self.e_dict(c1, context_labels, t1)["logits"],
                           "ce": self.e_dict(c2, context_labels, t2)["logits"],
                           "sup": self.supportKD(c2, context_labels, t2)["logits"]}}


class e_dist_1fc_sup(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.e_dict = e_dist(args)
        self.supportKD = SupportDK(args)

    def forward(self, context_feature, context_labels, target_feature):
        return {"logits": {"kl": self.e_dict(context_feature, context_labels, target_feature)["logits"],
                           "sup": self.supportKD(context_feature, context_labels, target_feature)["logits"]}}


class e_dist_fc2_sup_fixed(nn.Module):
    """Teacher-side variant; like the reference (:203-231) it does NOT disable gradients."""

    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.e_dict = e_dist(args)
        self.supportKD = SupportDK(args)

    def forward(self, context_feature, context_labels, target_feature):
        return {"logits": {"kl": self.e_dict(context_feature, context_labels, target_feature)["logits"],
                           "sup": self.supportKD(context_feature, context_labels, target_feature)["logits"]}}
