"""Oracle restatement of the student feature heads feeding the path (test infrastructure, see oracle/__init__.py).

SURVEY.md §8f rank 1.  Citations are reference paths (file:line)."""
from __future__ import annotations

import torch


def frame_pool(fmap: torch.Tensor, out_hw: int = 4) -> torch.Tensor:
    """model/backbone/resnet18_2fc.py:41-53 -- adaptive max pool to out_hw x out_hw, then the mean over the
    out_hw^2 patches (the reshape / permute in between only reorders the patches).  [R, C, H, W] -> [R, C]."""
    R, C, H, W = fmap.shape
    cells = []
    for i in range(out_hw):
        h0, h1 = (i * H) // out_hw, -((-(i + 1) * H) // out_hw)
        for j in range(out_hw):
            w0, w1 = (j * W) // out_hw, -((-(j + 1) * W) // out_hw)
            cells.append(fmap[:, :, h0:h1, w0:w1].amax(dim=(2, 3)))
    return torch.stack(cells, dim=-1).mean(dim=-1)


def feature_heads(pooled: torch.Tensor, weights, biases, seq_len: int):
    """model/backbone/resnet18_2fc.py:55-64 -- one Linear per head on the pooled frame features, reshaped to
    [videos, seq_len, out].  `weights` / `biases` are per-head lists; returns a list of tensors."""
    return [(pooled @ w.t() + b).reshape(-1, seq_len, w.shape[0]) for w, b in zip(weights, biases)]
