"""Diagnostic: batched TRX{2,3} call vs the same episodes run alone, and run-to-run repeatability (GPU)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lite-mkd_b200")]
import torch
import model.classifiers as C
from lmkd.episodes import make_episodes

d = torch.device("cuda:0")
torch.manual_seed(3)
args = types.SimpleNamespace(seq_len=8, trans_dropout=0.0, trans_linear_out_dim=1152, trans_linear_in_dim=2048,
                             way=5, shot=5, temp_set=[2, 3])
head = C.TrxBranch(args).to(d).eval()
ep = make_episodes(64, 5, 5, 5, 8, 2048, teacher_dim=8, device=d, seed=49)
up = torch.randn(64, 25, 5, device=d)


def run(sl):
    S = ep.support[sl].detach().clone().requires_grad_(True)
    Q = ep.query[sl].detach().clone().requires_grad_(True)
    lg = head(S, ep.support_labels[sl], Q)["logits"]
    for p in head.parameters():
        p.grad = None
    (lg * up[sl]).sum().backward()
    return lg.detach(), S.grad.clone(), Q.grad.clone()


def cmp(a, b):
    return f"max|d| {(a - b).abs().max().item():.3e}  rel-l2 {((a - b).norm() / b.norm()).item():.3e}  max|ref| {b.abs().max().item():.3e}"


lg, gs, gq = run(slice(0, 64))
lg2, gs2, gq2 = run(slice(0, 64))
print("repeat  logits", cmp(lg2, lg))
print("repeat  gS    ", cmp(gs2, gs))
print("repeat  gQ    ", cmp(gq2, gq))
for b in (0, 31, 63):
    l1, s1, q1 = run(slice(b, b + 1))
    print(f"b={b} logits", cmp(l1[0], lg[b]))
    print(f"b={b} gS    ", cmp(s1[0], gs[b]))
    print(f"b={b} gQ    ", cmp(q1[0], gq[b]))
    bad = ((s1[0] - gs[b]).abs() > 1e-5 + 1e-3 * gs[b].abs())
    print(f"b={b} violations of rtol 1e-3/atol 1e-5: {int(bad.sum())} of {bad.numel()}")
