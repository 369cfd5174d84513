"""Throughput of the teacher multi-modal fusion forward at the reference's shipped depth
(trans_num 4): MultiModalFusion.extract_feature on `videos` 8-frame videos, CUDA-event timed.
    python tools/fusion_bench.py [videos]"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import model.fusion as MF  # noqa: E402

dev = torch.device("cuda:0")
videos = int(sys.argv[1]) if len(sys.argv) > 1 else 1600
args = types.SimpleNamespace(seq_len=8, trans_linear_in_dim=2048, trans_num=4, shirt_num=1, num_gpus=1)
torch.manual_seed(0)
m = MF.MultiModalFusion(args).to(dev).eval()
feat = {k: torch.randn(videos, 8, 2048, device=dev) for k in ("rgb", "depth", "flow")}
for _ in range(3):
    out = m.extract_feature(feat)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
iters = 5
ev[0].record()
for _ in range(iters):
    out = m.extract_feature(feat)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / iters
M = videos * 8


def enc_flops(d, dff=2048, layers=4, dout=2048):
    return layers * (2 * M * d * 3 * d + 2 * M * d * d + 4 * M * d * dff) + 2 * M * d * dout


flops = enc_flops(6144) + 2 * enc_flops(4096)
print(json.dumps({"videos": videos, "ms": ms, "videos_per_s": videos / ms * 1e3, "linear_tflops": flops / ms / 1e9,
                  "finite": bool(torch.isfinite(out).all())}))
