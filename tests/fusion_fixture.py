"""Shared by tests/golden/make_golden.py (reference modules) and the fusion tests (oracle / product modules):
deterministic parameters for the teacher's fusion encoders.  The encoders are too large to store (the reference
hard-codes 2048 features per modality: d_model 6144 / 4096), so the fixture holds inputs and outputs only and both
sides regenerate the parameters from numpy's stream-stable RandomState, in state_dict key order."""
import types
import zlib

import numpy as np
import torch

SEED = 3483
TRANS_NUM = 2       # encoder layers in the fixture (the reference's default is 4; same code path per layer)
SHIFT = 1           # args.shirt_num default (teacher/code/extract_multi_feature.py:96)
N_VIDEOS = 6


def fusion_args():
    return types.SimpleNamespace(seq_len=8, trans_linear_in_dim=2048, trans_num=TRANS_NUM, shirt_num=SHIFT, num_gpus=1)


def fill_parameters(module: torch.nn.Module, tag: str) -> None:
    """Overwrite every parameter of `module` in place; values depend only on (tag, key name, shape)."""
    with torch.no_grad():
        for name, p in module.state_dict().items():
            rs = np.random.RandomState((SEED + zlib.crc32(f"{tag}/{name}".encode())) % (2 ** 31))
            if p.dim() == 2 and "position_embeddings" not in name:
                v = rs.standard_normal(p.shape).astype(np.float32) / np.sqrt(p.shape[1])
            elif p.dim() == 2:
                v = 0.5 * rs.standard_normal(p.shape).astype(np.float32)
            elif name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("LayerNorm.weight"):
                v = 1.0 + 0.2 * rs.standard_normal(p.shape).astype(np.float32)
            else:
                v = 0.1 * rs.standard_normal(p.shape).astype(np.float32)
            p.copy_(torch.from_numpy(v))


def modality_inputs():
    rs = np.random.RandomState(SEED + 77)
    return [rs.standard_normal((N_VIDEOS, 8, 2048)).astype(np.float32) for _ in range(3)]
