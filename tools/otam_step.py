"""One OTAM forward + backward at config 4 (B episodes) -- target for an ncu launch list.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/otam_step.py 4096
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from lmkd import ops  # noqa: E402
from lmkd.episodes import make_episodes  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
ep = make_episodes(B, 5, 5, 5, 8, 2048, teacher_dim=8, device=dev)
S, Q = ep.support.requires_grad_(True), ep.query.requires_grad_(True)
up = torch.randn(B, 25, 5, device=dev)
for _ in range(2):
    S.grad = Q.grad = None
    (ops.otam_probs(S, ep.support_labels, Q, 5) * up).sum().backward()
torch.cuda.synchronize()
print("ok")
