"""D2M teacher-to-student losses on the lmkd CUDA path.

Same public surface as the reference's distillers.py (`kd_loss`, `inter_class_relation`,
`class Distiller(distill_name, distill_cfg, device)` with one method per recipe, selected by
name with getattr — trainwandb.py:231), different machinery: every recipe is a short table of
(term kind, student logits, target, weight) rows that ONE kernel launch evaluates per episode,
values and gradients together (`lmkd_d2m_logit_loss`), plus one streaming pass for the feature
MSE (`lmkd_d2m_feature_mse_fwdbwd`).

Logits may be [rows, cols] (one episode, as the reference passes them) or [B, rows, cols]
(batched episodes; the returned 'loss' is then the sum over episodes, which is what the
reference's gradient accumulation over tasks_per_batch computes, trainwandb.py:141-143).
The literal /16 on the CE terms is the reference's (distillers.py:70 etc.), not tasks_per_batch.
"""
from __future__ import annotations

import torch

from lmkd import ops
from lmkd.ops import TERM_CE, TERM_ICR, TERM_KD

CE, KL, ICR = TERM_CE, TERM_KD, TERM_ICR


def _as3d(x):
    return x if x.dim() == 3 else x.unsqueeze(0)


def _run(terms, temperature, labels, focal=None):
    """terms: [(kind, student, target_or_None(labels), w, fa, fb)] -> (loss [B], values [B,n], focal [B])."""
    students, index = [], {}
    rows = []
    for kind, s, target, w, fa, fb in terms:
        if id(s) not in index:
            index[id(s)] = len(students)
            students.append(_as3d(s))
        tgt = labels if kind == CE else _as3d(target)
        rows.append((kind, index[id(s)], tgt, float(w), float(fa), float(fb)))
    B = students[0].shape[0]
    lab = labels.reshape(B, -1)
    rows = [(k, si, lab if k == CE else t, w, fa, fb) for (k, si, t, w, fa, fb) in rows]
    spec = {"terms": rows, "T": float(temperature), "B": B, "focal": None}
    if focal is not None:
        num, den = focal
        spec["focal"] = (index[id(num)], index[id(den)], lab)
    return ops.logit_loss(spec, students)


def kd_loss(logits_student, logits_teacher, temperature):
    """T^2 * mean_rows KL(softmax(t/T) || softmax(s/T))  (reference distillers.py:7-15)."""
    dummy = torch.zeros(_as3d(logits_student).shape[:2], dtype=torch.int64, device=logits_student.device)
    loss, _, _ = _run([(KL, logits_student, logits_teacher, 1.0, 1.0, 0.0)], temperature, dummy)
    return loss.sum()


def inter_class_relation(y_s, y_t):
    """1 - mean_rows pearson(softmax(y_s), softmax(y_t))  (reference distillers.py:26-30)."""
    dummy = torch.zeros(_as3d(y_s).shape[:2], dtype=torch.int64, device=y_s.device)
    loss, _, _ = _run([(ICR, y_s, y_t, 1.0, 1.0, 0.0)], 1.0, dummy)
    return loss.sum()


class Distiller(object):
    def __init__(self, distill_name, distill_cfg, device):
        self.distill_name = distill_name
        self.distill_dict = distill_cfg
        self.device = device

    # ---- helpers -----------------------------------------------------------------------------
    def _dev(self, x):
        if isinstance(x, dict):
            return {k: self._dev(v) for k, v in x.items()}
        return x.to(self.device) if torch.is_tensor(x) else x

    def _finish(self, terms, labels, groups, focal=None, extra=None):
        """Evaluate the table; `groups` maps report keys to [(term index, coefficient fn(focal))]."""
        cfg = self.distill_dict
        loss, values, fw = _run(terms, cfg["temperature"], labels, focal)
        out = {"loss": loss.sum() if extra is None else loss.sum() + extra}
        out["loss_per_episode"] = loss
        v, f = values.detach(), fw.detach()
        for key, parts in groups.items():
            out[key] = sum((c(f) if callable(c) else c) * v[:, i] for i, c in parts).sum()
        return out, f

    # ---- single-logit recipes ------------------------------------------------------------------
    def KD(self, student_logits, teacher_logits, test_labels):                    # reference :42-74
        c = self.distill_dict
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        hw, sw = c["hard_loss_weight"] / 16.0, c["soft_loss_weight"]
        out, _ = self._finish([(CE, s, None, hw, 1, 0), (KL, s, t, sw, 1, 0)], test_labels,
                              {"hard_loss": [(0, hw)], "soft_loss": [(1, sw)]})
        return out

    def wsl(self, student_logits, teacher_logits, test_labels):                    # :76-98
        c = self.distill_dict
        s, t = self._dev(student_logits), self._dev(teacher_logits).detach()
        sw, hw = c["soft_loss_weight"], c["hard_loss_weight"] / 16.0
        # focal = 1 - exp(-CE(s)/CE(t)); the teacher only enters through a zero-weight CE row
        terms = [(KL, s, t, sw, 0, 1), (CE, s, None, hw, 1, 0), (CE, t, None, 0.0, 0, 0)]
        out, _ = self._finish(terms, test_labels, {"soft_loss": [(0, lambda f: sw * f)], "hard_loss": [(1, hw)]},
                              focal=(s, t))
        return out

    def ce(self, student_logits, teacher_logits, test_labels):                     # :100-108
        s = self._dev(student_logits)
        out, _ = self._finish([(CE, s, None, 1 / 16.0, 1, 0)], test_labels, {})
        return out

    def Dist_KD(self, student_logits, teacher_logits, test_labels):                # :286-293
        c = self.distill_dict
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        hw, sw = c["hard_loss_weight"] / 16.0, c["soft_loss_weight"]
        out, _ = self._finish([(CE, s, None, hw, 1, 0), (ICR, s, t, sw, 1, 0)], test_labels,
                              {"hard_loss": [(0, hw)], "soft_loss": [(1, sw)]})
        return out

    def support_sim(self, student_logits, teacher_logits, test_labels):            # :110-124
        c = self.distill_dict
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        # the reference hard-codes reshape(20, 25): 20 queries x (5 x 5) prototype similarities
        ss = s["support_set"].reshape(*s["support_set"].shape[:-2], -1)
        ts = t["support_set"].reshape(*t["support_set"].shape[:-2], -1)
        hw = c["hard_loss_weight"] / 16.0
        terms = [(CE, s["query"], None, hw, 1, 0), (KL, ss, ts, c["soft_loss_weight_support"], 1, 0),
                 (KL, s["query"], t["query"], c["soft_loss_weight_query"], 1, 0)]
        out, _ = self._finish(terms, test_labels, {"hard_loss": [(0, hw)],
                                                   "soft_support_loss": [(1, c["soft_loss_weight_support"])],
                                                   "soft_query_loss": [(2, c["soft_loss_weight_query"])]})
        return out

    def KL_feature(self, student_logits, teacher_logits, test_labels):             # :126-150
        c = self.distill_dict
        s, t = self._dev(student_logits["logits"]), self._dev(teacher_logits["logits"])
        sf, tf = student_logits["feature"], teacher_logits["feature"]
        n_ep = sf.numel() // _as3d(s).shape[0]
        feat = ops.feature_mse(sf, tf, c["feature_loss_weight"], n_ep)
        hw, sw = c["hard_loss_weight"] / 16.0, c["soft_loss_weight"]
        out, _ = self._finish([(CE, s, None, hw, 1, 0), (KL, s, t, sw, 1, 0)], test_labels,
                              {"hard_loss": [(0, hw)], "soft_loss": [(1, sw)]}, extra=feat)
        out["feature_loss"] = feat.detach()
        return out

    # ---- two-head student, tensor teacher --------------------------------------------------------
    def fc_2(self, student_logits, teacher_logits, test_labels):                   # :152-161
        c = self.distill_dict
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        hw, sw = c["hard_loss_weight"] / 16.0, c["soft_loss_weight"]
        out, _ = self._finish([(CE, s["fc_1"], None, hw, 1, 0), (KL, s["fc_2"], t, sw, 1, 0)], test_labels,
                              {"hard_loss": [(0, hw)], "soft_loss": [(1, sw)]})
        return out

    def fc_2_wsl(self, student_logits, teacher_logits, test_labels):               # :163-201
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        terms = [(KL, s["fc_2"], t, 1.0, 1, 1), (CE, s["fc_1"], None, 1 / 16.0, 2, -1)]
        out, f = self._finish(terms, test_labels, {"soft_loss": [(0, lambda f: 1 + f)],
                                                   "hard_loss": [(1, lambda f: (2 - f) / 16.0)]},
                              focal=(s["fc_1"], s["fc_2"]))
        out["aerfa"] = f
        self.distill_dict["fcwsl_aerfa"] = f
        return out

    def strm(self, student_logits, teacher_logits, test_labels):                   # :203-213
        s = self._dev(student_logits)
        out, _ = self._finish([(CE, s["pat"], None, 0.1 / 16, 1, 0), (CE, s["fr"], None, 1 / 16.0, 1, 0)], test_labels,
                              {"pat_loss": [(0, 1 / 16.0)], "fr_loss": [(1, 1 / 16.0)]})
        return out

    def strm_KD(self, student_logits, teacher_logits, test_labels):                # :215-227
        c = self.distill_dict
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        terms = [(CE, s["pat"], None, 0.1 / 16, 1, 0), (CE, s["fr"], None, 1 / 16.0, 1, 0),
                 (KL, s["fr"], t, c["soft_loss_weight"], 1, 0)]
        out, _ = self._finish(terms, test_labels, {"pat_loss": [(0, 1 / 16.0)], "fr_loss": [(1, 1 / 16.0)],
                                                   "softloss": [(2, c["soft_loss_weight"])]})
        return out

    # ---- {'kl','ce','sup'} student vs {'kl','sup'} teacher -----------------------------------------
    def fc_2_sup(self, student_logits, teacher_logits, test_labels):               # :229-284
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        terms = [(KL, s["kl"], t["kl"], 1.0, 1, 1), (KL, s["sup"], t["sup"], 0.1 / 16, 2, -1),
                 (CE, s["ce"], None, 1 / 16.0, 2, -1)]
        out, _ = self._finish(terms, test_labels, {"soft_loss": [(0, 1.0)], "hard_loss": [(1, 0.01 / 16), (2, 1 / 16.0)]},
                              focal=(s["ce"], s["kl"]))
        return out

    def _klcesup(self, s, t, y, kl_kind, sup_kind, sup_w, extra_ce_on_kl=False, kl_w=1.0):
        terms = [(kl_kind, s["kl"], t["kl"], kl_w, 1, 0), (sup_kind, s["sup"], t["sup"], sup_w, 1, 0),
                 (CE, s["ce"], None, 1 / 16.0, 1, 0)]
        if extra_ce_on_kl:
            terms.append((CE, s["kl"], None, 1 / 16.0, 1, 0))
        out, _ = self._finish(terms, y, {"soft_loss": [(0, 1.0)], "hard_loss": [(1, sup_w), (2, 1 / 16.0)]})
        return out

    def fc_2_sup_dist(self, student_logits, teacher_logits, test_labels):          # :295-337 (shipped default)
        return self._klcesup(self._dev(student_logits), self._dev(teacher_logits), test_labels, KL, ICR, 0.5)

    def fc_2_sup_kl(self, student_logits, teacher_logits, test_labels):            # :339-383
        return self._klcesup(self._dev(student_logits), self._dev(teacher_logits), test_labels, KL, KL, 0.5)

    def fc_2_sup_dist_cece(self, student_logits, teacher_logits, test_labels):     # :385-429
        return self._klcesup(self._dev(student_logits), self._dev(teacher_logits), test_labels, KL, ICR, 0.5, True)

    def fc_2_sup_klklcece(self, student_logits, teacher_logits, test_labels):      # :431-475
        return self._klcesup(self._dev(student_logits), self._dev(teacher_logits), test_labels, KL, KL, 0.5, True)

    def fc_2_sup_distdistcece(self, student_logits, teacher_logits, test_labels):  # :477-499
        return self._klcesup(self._dev(student_logits), self._dev(teacher_logits), test_labels, ICR, ICR, 0.5, True)

    def fc_2_sup_2(self, student_logits, teacher_logits, test_labels):             # :501-547
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        terms = [(KL, s["kl"], t["kl"], 1.0, 1, 0), (ICR, s["sup_kl"], t["sup"], 1.0, 1, 0),
                 (CE, s["ce"], None, 1 / 16.0, 1, 0), (ICR, s["sup_ce"], t["sup"], 1.0, 1, 0)]
        out, _ = self._finish(terms, test_labels, {"soft_loss": [(0, 1.0), (1, 0.5)], "hard_loss": [(2, 1 / 16.0), (3, 0.5)]})
        return out

    def fc_2_sup_disver(self, student_logits, teacher_logits, test_labels):        # :549-572
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        terms = [(KL, s["sup"], t["sup"], 0.5, 1, 0), (ICR, s["kl"], t["kl"], 1.0, 1, 0),
                 (CE, s["ce"], None, 1 / 16.0, 1, 0), (CE, s["kl"], None, 1 / 16.0, 1, 0)]
        out, _ = self._finish(terms, test_labels, {"soft_loss": [(0, 1.0)], "hard_loss": [(1, 1.0), (2, 1 / 16.0)]})
        return out

    def fc_2_sup_dist_wsl(self, student_logits, teacher_logits, test_labels):      # :574-624
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        terms = [(KL, s["kl"], t["kl"], 1.0, 0.5, 1), (ICR, s["sup"], t["sup"], 0.5, 1.5, -1),
                 (CE, s["ce"], None, 1 / 16.0, 1.5, -1)]
        out, _ = self._finish(terms, test_labels, {"soft_loss": [(0, 1.0)], "hard_loss": [(1, 0.5), (2, 1 / 16.0)]},
                              focal=(s["ce"], s["kl"]))
        return out

    def _strm_sup(self, s, t, y, fr_kl_key, fr_ce_key):
        terms = [(KL, s[fr_kl_key], t["kl"], 1.0, 1, 0), (ICR, s["sup"], t["sup"], 0.5, 1, 0),
                 (CE, s[fr_ce_key], None, 1 / 16.0, 1, 0), (KL, s["pat"], t["kl"], 0.1, 1, 0),
                 (CE, s["pat"], None, 0.1 / 16, 1, 0)]
        out, _ = self._finish(terms, y, {})
        return out

    def strm_fc_2_sup_dist(self, student_logits, teacher_logits, test_labels):     # :626-653
        return self._strm_sup(self._dev(student_logits), self._dev(teacher_logits), test_labels, "fr1", "fr2")

    def strm_1fc_sup(self, student_logits, teacher_logits, test_labels):           # :655-681
        return self._strm_sup(self._dev(student_logits), self._dev(teacher_logits), test_labels, "fr", "fr")

    def fc_1_sup(self, student_logits, teacher_logits, test_labels):               # :683-696
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        terms = [(CE, s["kl"], None, 1 / 16.0, 1, 0), (KL, s["kl"], t["kl"], 1.0, 1, 0),
                 (ICR, s["sup"], t["sup"], 0.5, 1, 0)]
        out, _ = self._finish(terms, test_labels, {})
        return out

    def fc_sup(self, student_logits, teacher_logits, test_labels):                 # :698-711
        s, t = self._dev(student_logits), self._dev(teacher_logits)
        out, _ = self._finish([(CE, s["kl"], None, 1 / 16.0, 1, 0), (ICR, s["sup"], t["sup"], 0.5, 1, 0)],
                              test_labels, {})
        return out

    def e_dist_1fc_sup(self, student_logits, teacher_logits, test_labels):         # :713-733
        return self.fc_1_sup(student_logits, teacher_logits, test_labels)
