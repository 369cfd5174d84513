"""Oracle restatement of the teacher's multi-modal fusion forward (MFM feature extraction) -- test infrastructure,
see oracle/__init__.py.  SURVEY.md §8f rank 4.  Citations are reference paths (file:line).

The reference builds the encoders from torch's own nn.TransformerEncoderLayer with default arguments
(teacher/code/model.py:1313-1316, 1372-1375): post-norm (norm_first=False), ReLU, dim_feedforward=2048,
layer_norm_eps=1e-5, batch_first=True; extraction runs in eval() (teacher/code/extract_multi_feature.py:114), so every
dropout is the identity.  Parameters are passed as the modules' own state_dict (same key names), so the functions
below can be fed from the reference modules and from the product modules alike."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def trainable_positional_encoding(x, sd, prefix, eps=1e-5):
    """teacher/code/model.py:1135-1151 -- LayerNorm(x + position_embeddings[0..L)) (dropout: identity in eval)."""
    L = x.shape[1]
    emb = sd[prefix + "position_embeddings.weight"][:L]
    return F.layer_norm(x + emb, (x.shape[-1],), sd[prefix + "LayerNorm.weight"], sd[prefix + "LayerNorm.bias"], eps)


def encoder_layer(x, sd, prefix, nhead, eps=1e-5):
    """torch.nn.TransformerEncoderLayer forward, post-norm: x = LN1(x + SA(x)); x = LN2(x + W2 relu(W1 x))."""
    N, L, d = x.shape
    dh = d // nhead
    qkv = x @ sd[prefix + "self_attn.in_proj_weight"].t() + sd[prefix + "self_attn.in_proj_bias"]
    q, k, v = (t.reshape(N, L, nhead, dh).transpose(1, 2) for t in qkv.split(d, dim=-1))
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ctx = (att @ v).transpose(1, 2).reshape(N, L, d)
    sa = ctx @ sd[prefix + "self_attn.out_proj.weight"].t() + sd[prefix + "self_attn.out_proj.bias"]
    x = F.layer_norm(x + sa, (d,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps)
    ff = torch.relu(x @ sd[prefix + "linear1.weight"].t() + sd[prefix + "linear1.bias"])
    ff = ff @ sd[prefix + "linear2.weight"].t() + sd[prefix + "linear2.bias"]
    return F.layer_norm(x + ff, (d,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps)


def fusion_encoder(xs, sd, nhead, num_layers):
    """ThreeTransforTemproal.extract_feature / TwoTransforFusion.extract_feature (teacher/code/model.py:1385-1392,
    1325-1331): per-modality trainable positional encoding, concatenation along the feature axis, the encoder
    stack, then `f1` Linear down to 2048.  `xs`: list of [N, L, 2048] tensors, one per modality."""
    enc = [trainable_positional_encoding(x, sd, f"positionEncoding{i + 1}.") for i, x in enumerate(xs)]
    h = torch.cat(enc, dim=-1)
    for layer in range(num_layers):
        h = encoder_layer(h, sd, f"transformer_encoder.layers.{layer}.", nhead)
    return h @ sd["f1.weight"].t() + sd["f1.bias"]


def mfm_extract_feature(rgb, depth, flow, three_sd, two_sd, num_layers, shift):
    """ThreeTRXShiftLoopTime.extract_feature (teacher/code/model.py:1648-1664): the three-modality encoder on
    (rgb, depth, flow) plus the two-modality encoder on (rgb, depth rolled by `shift` frames) and on
    (rgb, flow rolled by `shift` frames), summed.  Inputs [N, L, 2048] -> [N, L, 2048]."""
    roll = lambda x: torch.cat((x[:, shift:], x[:, :shift]), dim=1)
    f1 = fusion_encoder([rgb, depth, flow], three_sd, 3, num_layers)
    f2 = fusion_encoder([rgb, roll(depth)], two_sd, 2, num_layers)
    f3 = fusion_encoder([rgb, roll(flow)], two_sd, 2, num_layers)
    return f1 + f2 + f3
