"""One config-2 TRX{2,3} student step (64 episodes, forward + backward) for `ncu -k regex:<kernel>` captures."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lite-mkd_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import model.classifiers as C  # noqa: E402
from lmkd.episodes import make_episodes  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
args = types.SimpleNamespace(seq_len=8, trans_dropout=0.1, trans_linear_out_dim=1152, trans_linear_in_dim=2048,
                             way=5, shot=5, temp_set=[2, 3])
head = C.TrxBranch(args).to(dev).train()
ep = make_episodes(B, 5, 5, 5, 8, 2048, teacher_dim=8, device=dev)
for _ in range(2):
    S, Q = ep.support.clone().requires_grad_(True), ep.query.clone().requires_grad_(True)
    head(S, ep.support_labels, Q)["logits"].square().sum().backward()
torch.cuda.synchronize()
print("ok")
