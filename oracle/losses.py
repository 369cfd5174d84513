"""Oracle restatement of the D2M distillation losses (test infrastructure, see oracle/__init__.py).

Follows distillers.py of the reference: module functions :7-30 and the 24 `Distiller`
recipes :42-733.  Written term-by-term so every recipe is a sum of four primitives
(CE, temperature KL, inter-class relation, feature MSE) with the reference's literal weights.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def kd_loss(s: torch.Tensor, t: torch.Tensor, temperature: float) -> torch.Tensor:
    """distillers.py:7-15 — T²·mean_rows Σ_c p_t (log p_t − log_softmax(s/T))."""
    log_ps = torch.log_softmax(s / temperature, dim=1)
    pt = torch.softmax(t / temperature, dim=1)
    log_pt = torch.log_softmax(t / temperature, dim=1)
    return (pt * (log_pt - log_ps)).sum(dim=1).mean() * temperature ** 2


def _pearson(x, y, eps=1e-8):
    """distillers.py:18-23."""
    x = x - x.mean(dim=1, keepdim=True)
    y = y - y.mean(dim=1, keepdim=True)
    return (x * y).sum(dim=1) / (x.norm(dim=1) * y.norm(dim=1) + eps)


def inter_class_relation(ys: torch.Tensor, yt: torch.Tensor) -> torch.Tensor:
    """distillers.py:26-30 — 1 − mean_rows Pearson(softmax(ys), softmax(yt))."""
    return 1.0 - _pearson(torch.softmax(ys, dim=1), torch.softmax(yt, dim=1)).mean()


def cross_entropy(s: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return F.cross_entropy(s, y)


def mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return ((a - b) ** 2).mean()


def aggregate_accuracy(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """utils.py:116-121."""
    return (labels == logits.argmax(dim=-1)).float().mean()


def _focal(num: torch.Tensor, den: torch.Tensor) -> torch.Tensor:
    """1 − exp(−max(num/(den+1e-8), 0)) on detached CE values (distillers.py:86-93)."""
    w = num.detach() / (den.detach() + 1e-8)
    return 1.0 - torch.exp(-torch.clamp(w, min=0.0))


class Recipes:
    """One method per `Distiller` recipe; `cfg` is the --cfg dict (options.py:51-60)."""

    def __init__(self, cfg: dict):
        self.cfg = cfg
        self.T = cfg["temperature"]

    # -- single-logit recipes --------------------------------------------------------
    def KD(self, s, t, y):                                   # distillers.py:42-74
        return self.cfg["hard_loss_weight"] * cross_entropy(s, y) / 16 \
            + self.cfg["soft_loss_weight"] * kd_loss(s, t, self.T)

    def wsl(self, s, t, y):                                  # :76-98
        fw = _focal(cross_entropy(s, y), cross_entropy(t, y))
        return self.cfg["soft_loss_weight"] * fw * kd_loss(s, t, self.T) \
            + self.cfg["hard_loss_weight"] * cross_entropy(s, y) / 16

    def ce(self, s, t, y):                                   # :100-108
        return cross_entropy(s, y) / 16

    def Dist_KD(self, s, t, y):                              # :286-293
        return self.cfg["hard_loss_weight"] * cross_entropy(s, y) / 16 \
            + self.cfg["soft_loss_weight"] * inter_class_relation(s, t)

    def support_sim(self, s, t, y):                          # :110-124
        c = self.cfg
        return c["hard_loss_weight"] * cross_entropy(s["query"], y) / 16 \
            + c["soft_loss_weight_support"] * kd_loss(s["support_set"].reshape(20, 25),
                                                      t["support_set"].reshape(20, 25), self.T) \
            + c["soft_loss_weight_query"] * kd_loss(s["query"], t["query"], self.T)

    def KL_feature(self, s, t, y):                           # :126-150
        c = self.cfg
        return c["hard_loss_weight"] * cross_entropy(s["logits"], y) / 16 \
            + c["soft_loss_weight"] * kd_loss(s["logits"], t["logits"], self.T) \
            + c["feature_loss_weight"] * mse(s["feature"], t["feature"])

    # -- two-head student, tensor teacher ---------------------------------------------
    def fc_2(self, s, t, y):                                 # :152-161
        return self.cfg["hard_loss_weight"] * cross_entropy(s["fc_1"], y) / 16 \
            + self.cfg["soft_loss_weight"] * kd_loss(s["fc_2"], t, self.T)

    def fc_2_wsl(self, s, t, y):                             # :163-201
        fw = _focal(cross_entropy(s["fc_1"], y), cross_entropy(s["fc_2"], y))
        return (1 + fw) * kd_loss(s["fc_2"], t, self.T) + (2 - fw) * cross_entropy(s["fc_1"], y) / 16

    def strm(self, s, t, y):                                 # :203-213
        return 0.1 * cross_entropy(s["pat"], y) / 16 + cross_entropy(s["fr"], y) / 16

    def strm_KD(self, s, t, y):                              # :215-227
        return 0.1 * cross_entropy(s["pat"], y) / 16 + cross_entropy(s["fr"], y) / 16 \
            + self.cfg["soft_loss_weight"] * kd_loss(s["fr"], t, self.T)

    # -- {'kl','ce','sup'} student vs {'kl','sup'} teacher ----------------------------
    def fc_2_sup(self, s, t, y):                             # :229-284
        fw = _focal(cross_entropy(s["ce"], y), cross_entropy(s["kl"], y))
        kl = kd_loss(s["kl"], t["kl"], self.T)
        sup = kd_loss(s["sup"], t["sup"], self.T) / 16
        ce = cross_entropy(s["ce"], y) / 16
        return (1 + fw) * kl + (2 - fw) * (0.1 * sup + ce)

    def fc_2_sup_dist(self, s, t, y):                        # :295-337 (shipped default)
        return kd_loss(s["kl"], t["kl"], self.T) + 0.5 * inter_class_relation(s["sup"], t["sup"]) \
            + cross_entropy(s["ce"], y) / 16

    def fc_2_sup_kl(self, s, t, y):                          # :339-383
        return kd_loss(s["kl"], t["kl"], self.T) + 0.5 * kd_loss(s["sup"], t["sup"], self.T) \
            + cross_entropy(s["ce"], y) / 16

    def fc_2_sup_dist_cece(self, s, t, y):                   # :385-429
        return self.fc_2_sup_dist(s, t, y) + cross_entropy(s["kl"], y) / 16

    def fc_2_sup_klklcece(self, s, t, y):                    # :431-475
        return self.fc_2_sup_kl(s, t, y) + cross_entropy(s["kl"], y) / 16

    def fc_2_sup_distdistcece(self, s, t, y):                # :477-499
        return inter_class_relation(s["kl"], t["kl"]) + cross_entropy(s["kl"], y) / 16 \
            + 0.5 * inter_class_relation(s["sup"], t["sup"]) + cross_entropy(s["ce"], y) / 16

    def fc_2_sup_2(self, s, t, y):                           # :501-547
        return kd_loss(s["kl"], t["kl"], self.T) + inter_class_relation(s["sup_kl"], t["sup"]) \
            + cross_entropy(s["ce"], y) / 16 + inter_class_relation(s["sup_ce"], t["sup"])

    def fc_2_sup_disver(self, s, t, y):                      # :549-572
        return 0.5 * kd_loss(s["sup"], t["sup"], self.T) + inter_class_relation(s["kl"], t["kl"]) \
            + cross_entropy(s["ce"], y) / 16 + cross_entropy(s["kl"], y) / 16

    def fc_2_sup_dist_wsl(self, s, t, y):                    # :574-624
        fw = _focal(cross_entropy(s["ce"], y), cross_entropy(s["kl"], y))
        kl = kd_loss(s["kl"], t["kl"], self.T)
        hard = 0.5 * inter_class_relation(s["sup"], t["sup"]) + cross_entropy(s["ce"], y) / 16
        return (0.5 + fw) * kl + (1.5 - fw) * hard

    def strm_fc_2_sup_dist(self, s, t, y):                   # :626-653
        return kd_loss(s["fr1"], t["kl"], self.T) + 0.5 * inter_class_relation(s["sup"], t["sup"]) \
            + cross_entropy(s["fr2"], y) / 16 \
            + 0.1 * (kd_loss(s["pat"], t["kl"], self.T) + cross_entropy(s["pat"], y) / 16)

    def strm_1fc_sup(self, s, t, y):                         # :655-681
        return kd_loss(s["fr"], t["kl"], self.T) + 0.5 * inter_class_relation(s["sup"], t["sup"]) \
            + cross_entropy(s["fr"], y) / 16 \
            + 0.1 * (kd_loss(s["pat"], t["kl"], self.T) + cross_entropy(s["pat"], y) / 16)

    def fc_1_sup(self, s, t, y):                             # :683-696
        return cross_entropy(s["kl"], y) / 16 + kd_loss(s["kl"], t["kl"], self.T) \
            + 0.5 * inter_class_relation(s["sup"], t["sup"])

    def fc_sup(self, s, t, y):                               # :698-711
        return cross_entropy(s["kl"], y) / 16 + 0.5 * inter_class_relation(s["sup"], t["sup"])

    def e_dist_1fc_sup(self, s, t, y):                       # :713-733
        return self.fc_1_sup(s, t, y)


RECIPE_NAMES = [n for n in vars(Recipes) if not n.startswith("_")]
