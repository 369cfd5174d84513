"""OTAM head behind the classifier API.

The reference keeps OTAM on the teacher side only (teacher/code/model.py:3312-3343 CNN_OTAM);
this exposes the same computation as `model.classifiers.OTAM(args)` so `--model_classifier OTAM`
resolves.  Output = class probabilities softmax(-class mean distance), as CNN_OTAM returns.
"""
import torch.nn as nn

from lmkd import ops


class OTAM(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.lbda = float(getattr(args, "otam_lambda", 0.1))     # OTAM_cum_dist default (model.py:3271)
        self.eps = float(getattr(args, "otam_eps", 0.01))        # cos_sim default (model.py:3260)

    def forward(self, context_feature, context_labels, target_feature):
        way = int(self.args.way)
        if context_feature.dim() == 4:
            return {"logits": ops.otam_probs(context_feature, context_labels, target_feature, way, self.lbda, self.eps)}
        probs = ops.otam_probs(context_feature.unsqueeze(0), context_labels.reshape(1, -1),
                               target_feature.unsqueeze(0), way, self.lbda, self.eps)
        return {"logits": probs[0]}


CNN_OTAM = OTAM
