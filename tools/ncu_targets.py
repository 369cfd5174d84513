"""Small, fixed workload for `ncu --set full`: one launch each of the kernels the north_star asks evidence for.
    ncu --set full -k regex:"feat_mse|otam_dp|gemm_tcgen05" ... python tools/ncu_targets.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from lmkd import ops  # noqa: E402
from lmkd.episodes import make_episodes  # noqa: E402

dev = torch.device("cuda:0")
# fused D2M feature loss: 256 episodes of config 3 (2.5 GB moved)
n = 256 * 50 * 8 * 2048
s = torch.randn(n, device=dev).requires_grad_(True)
t = torch.randn(n, device=dev)
ops.feature_mse(s, t, 1.0, 50 * 8 * 2048).backward()
del s, t
# OTAM: 512 episodes of config 4 (similarity GEMM + wavefront DP fwd/bwd)
ep = make_episodes(512, 5, 5, 5, 8, 2048, teacher_dim=8, device=dev)
S, Q = ep.support.requires_grad_(True), ep.query.requires_grad_(True)
ops.otam_probs(S, ep.support_labels, Q, 5).sum().backward()
del ep, S, Q
# tcgen05 GEMM: projection (CTA pairs), scores (single CTA, TMA-store epilogue)
for (M, N, K, nb, amn, bmn) in [(25600, 6912, 2048, 1, 0, 0), (1400, 1440, 1152, 64, 0, 0)]:
    A = torch.randn((nb, K, M) if amn else (nb, M, K), device=dev).bfloat16()
    B = torch.randn((nb, K, N) if bmn else (nb, N, K), device=dev).bfloat16()
    ops.gemm_bf16(A, B, a_mn=bool(amn), b_mn=bool(bmn))
    del A, B
torch.cuda.synchronize()
print("ok")
