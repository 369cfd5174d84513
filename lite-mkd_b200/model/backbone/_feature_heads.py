"""Patch pooling + Linear feature heads between the CNN trunk and the matching heads.

The trunk itself (ResNet) is out of scope and stays stock PyTorch; what follows it in the reference --
AdaptiveMaxPool2d((4, 4)), mean over the 16 patches, one or two Linear(512 -> 2048) layers, reshape to
[videos, seq_len, 2048] (model/backbone/resnet18_2fc.py:41-67, resnet18_student.py:38-58) -- runs in liblmkd:
one pooling kernel and tcgen05 GEMMs with a bias epilogue, supports and queries in one launch.
"""
import torch
import torch.nn as nn

from lmkd import ops


def make_trunk(name: str = "resnet18"):
    """Stock torchvision trunk without its pooling / classifier layers (the reference's `self.resnet`).
    Weights come from the checkpoint the caller loads; nothing is downloaded."""
    import torchvision.models as models
    net = getattr(models, name)(weights=None)
    return nn.Sequential(*list(net.children())[:-2])


class PooledLinearHeads(nn.Module):
    """Holds the reference's Linear layers under the reference's attribute names (`names`), applies them all
    to the pooled frame features of supports and queries together."""

    def __init__(self, names, in_dim: int = 512, out_dim: int = 2048, out_hw: int = 4):
        super().__init__()
        self.names = tuple(names)
        self.out_hw = out_hw
        self.layers = nn.ModuleDict({n: nn.Linear(in_dim, out_dim) for n in self.names})

    def pooled(self, fmap: torch.Tensor) -> torch.Tensor:
        if fmap.dim() == 2:                       # already pooled [rows, in_dim]
            return fmap
        return ops.frame_pool(fmap, self.out_hw)

    def forward(self, context_maps, target_maps, seq_len: int):
        pc, pt = self.pooled(context_maps), self.pooled(target_maps)
        x = torch.cat([pc, pt], dim=0)
        w = torch.stack([self.layers[n].weight for n in self.names])
        b = torch.stack([self.layers[n].bias for n in self.names])
        y = ops.feature_heads(x, w, b)            # [heads, rows, out]
        nc = pc.shape[0]
        out_dim = y.shape[-1]
        ctx = [y[h, :nc].reshape(-1, seq_len, out_dim) for h in range(len(self.names))]
        tgt = [y[h, nc:].reshape(-1, seq_len, out_dim) for h in range(len(self.names))]
        return ctx, tgt
