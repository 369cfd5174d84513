"""One STRM DistanceLoss forward + backward at the config-2 episode shape (64 episodes, 8 x 2048-d, pairs) -- target
for an ncu launch list.    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/strm_step.py"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import model.classifiers as C  # noqa: E402
from lmkd.episodes import make_episodes  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
args = types.SimpleNamespace(seq_len=8, trans_dropout=0.1, trans_linear_out_dim=1152, trans_linear_in_dim=2048,
                             way=5, shot=5, device="cuda:0")
head = C.DistanceLoss(args, 2).to(dev).train()
ep = make_episodes(B, 5, 5, 5, 8, 2048, teacher_dim=8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(3):
    if it == 2:
        ev[0].record()
    S, Q = ep.support.clone().requires_grad_(True), ep.query.clone().requires_grad_(True)
    out = head.forward_batched(S, ep.support_labels, Q)
    out = out["logits"] if isinstance(out, dict) else out
    out.square().sum().backward()
ev[1].record()
torch.cuda.synchronize()
print("ok", B, "episodes fwd+bwd ms", ev[0].elapsed_time(ev[1]))
