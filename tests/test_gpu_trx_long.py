"""TRX on shapes whose class groups do not fit tensor memory (BASELINE config 5: 32 frames, triples -> T = 4960
tuples, 24800 support tuples per class): the materialised path with scores / probabilities produced for a few
queries at a time (lmkd_trx_set_attn_budget), the backward recomputing them per pass.

Parity is against the fp32 oracle restatement of TemporalCrossTransformer (oracle/matching.py, pinned to the
reference by tests/golden) at reduced query count / feature dim -- the reference itself cannot run this shape
(246 GB of scores per episode) -- plus properties at the full config-5 shape.
Tolerances as in test_gpu_trx.py: logits 1e-2 relative to the logit scale, gradients rel-L2 (stated per case).
"""
import types

import pytest
import torch

from conftest import assert_close, record_error, rel_l2

pytestmark = pytest.mark.gpu


def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _head(L, D, dout, way, shot, card, d):
    import model.classifiers as C
    torch.manual_seed(5)
    args = types.SimpleNamespace(seq_len=L, trans_dropout=0.0, trans_linear_out_dim=dout, trans_linear_in_dim=D,
                                 way=way, shot=shot, temp_set=[card])
    branch = C.TrxBranch(args).eval()
    with torch.no_grad():
        for m in branch.transformers:
            m.norm_k.weight.uniform_(0.5, 1.5)
            m.norm_k.bias.uniform_(-0.1, 0.1)
    m = branch.transformers[0]
    heads = [dict(Wk=m.k_linear.weight.detach().clone().requires_grad_(True),
                  bk=m.k_linear.bias.detach().clone().requires_grad_(True),
                  Wv=m.v_linear.weight.detach().clone().requires_grad_(True),
                  bv=m.v_linear.bias.detach().clone().requires_grad_(True),
                  gk=m.norm_k.weight.detach().clone().requires_grad_(True),
                  bek=m.norm_k.bias.detach().clone().requires_grad_(True), card=card)]
    return branch.to(d), heads


@pytest.mark.parametrize("budget", [0.0, 1.0])      # 0 = default (one pass), 1 byte = one query per pass
def test_query_chunked_passes_match_oracle(budget):
    """A small long-clip shape (12 frames, triples: KTp = 448 > 384 -> materialised path) run in one pass and with
    the budget forcing one query per pass: both must match the oracle, forward and every gradient."""
    import oracle
    from lmkd._ffi import lib
    from lmkd.episodes import make_episodes
    d = dev()
    B, way, shot, qpc, L, D, dout, card = 2, 3, 2, 2, 12, 128, 64, 3
    branch, heads = _head(L, D, dout, way, shot, card, d)
    ep = make_episodes(B, way, shot, qpc, L, D, teacher_dim=D, seed=21)
    up = torch.randn(B, way * qpc, way, generator=torch.Generator().manual_seed(9))
    lib().lmkd_trx_set_attn_budget(budget)
    try:
        S, Q = ep.support.to(d).requires_grad_(True), ep.query.to(d).requires_grad_(True)
        out = branch(S, ep.support_labels.to(d), Q)["logits"]
        (out * up.to(d)).sum().backward()
        torch.cuda.synchronize()
    finally:
        lib().lmkd_trx_set_attn_budget(0.0)
    gs, gq = [], []
    for b in range(B):
        s, q = ep.support[b].clone().requires_grad_(True), ep.query[b].clone().requires_grad_(True)
        ref = oracle.trx_branch_logits(s, ep.support_labels[b], q, heads, way)
        (ref * up[b]).sum().backward()
        gs.append(s.grad), gq.append(q.grad)
        assert_close(out[b], ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item(), what=f"logits_b{b}")
        assert (out[b].argmax(1).cpu() == ref.argmax(1)).all()
    m, h = branch.transformers[0], heads[0]
    assert rel_l2(S.grad, torch.stack(gs), "grad_support") < 1e-2
    assert rel_l2(Q.grad, torch.stack(gq), "grad_query") < 1e-2
    assert rel_l2(m.k_linear.weight.grad, h["Wk"].grad, "gWk") < 1e-2
    assert rel_l2(m.v_linear.weight.grad, h["Wv"].grad, "gWv") < 1e-2
    assert rel_l2(m.norm_k.weight.grad, h["gk"].grad, "ggamma") < 1e-2


def test_cfg5_triples_vs_oracle_reduced_queries():
    """Config 5 class group at full width (5 shots x 4960 triples of 32 frames = 24800 support tuples per class,
    d = 1152), 5 classes, ONE query, D = 256: logits and every gradient against the oracle.  (One query keeps the
    fp32 CPU oracle at ~9 TFLOP; the kernels are the ones the full shape uses.)"""
    import oracle
    from lmkd.episodes import make_episodes
    d = dev()
    way, shot, L, D, dout, card = 5, 5, 32, 256, 1152, 3
    branch, heads = _head(L, D, dout, way, shot, card, d)
    ep = make_episodes(1, way, shot, 1, L, D, teacher_dim=8, seed=33)
    qsel = torch.tensor([2])                                  # one of the `way` queries
    up = torch.randn(1, 1, way, generator=torch.Generator().manual_seed(4))
    S = ep.support.to(d).requires_grad_(True)
    Q = ep.query[:, qsel].contiguous().to(d).requires_grad_(True)
    out = branch(S, ep.support_labels.to(d), Q)["logits"]
    (out * up.to(d)).sum().backward()
    torch.cuda.synchronize()
    s = ep.support[0].clone().requires_grad_(True)
    q = ep.query[0, qsel].clone().requires_grad_(True)
    ref = oracle.trx_branch_logits(s, ep.support_labels[0], q, heads, way)
    (ref * up[0]).sum().backward()
    assert_close(out[0], ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item(), what="logits")
    assert (out[0].argmax(1).cpu() == ref.argmax(1)).all()
    assert int(out[0].argmax(1)) == int(ep.query_labels[0, qsel])
    m, h = branch.transformers[0], heads[0]
    # 4960 tuples per clip: ~150x more bf16 products per gradient element than at 8 frames; stated bound 3e-2
    assert rel_l2(S.grad[0], s.grad, "grad_support") < 3e-2
    assert rel_l2(Q.grad[0], q.grad, "grad_query") < 3e-2
    assert rel_l2(m.k_linear.weight.grad, h["Wk"].grad, "gWk") < 3e-2
    assert rel_l2(m.v_linear.weight.grad, h["Wv"].grad, "gWv") < 3e-2
    assert rel_l2(m.norm_k.weight.grad, h["gk"].grad, "ggamma") < 3e-2


def test_cfg5_full_shape_triples_forward_backward():
    """BASELINE config 5 at full shape for cardinality 3: 10-way 5-shot, 50 queries, 32 frames x 2048-d, d = 1152
    (283 TFLOP of attention per forward).  One query per pass inside the default 24 GB budget.  Properties:
    finite, every structured query classified correctly, a single-query call reproduces its rows, gradients
    reach features and head parameters."""
    import model.classifiers as C
    from lmkd.episodes import make_episodes
    d = dev()
    torch.manual_seed(0)
    args = types.SimpleNamespace(seq_len=32, trans_dropout=0.0, trans_linear_out_dim=1152, trans_linear_in_dim=2048,
                                 way=10, shot=5, temp_set=[3])
    head = C.TrxBranch(args).to(d).eval()
    ep = make_episodes(1, 10, 5, 5, 32, 2048, teacher_dim=8, device=d, seed=47)
    S, Q = ep.support.requires_grad_(True), ep.query.requires_grad_(True)
    t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0[0].record()
    lg = head(S, ep.support_labels, Q)["logits"]
    t0[1].record()
    lg.sum().backward()
    t0[2].record()
    torch.cuda.synchronize()
    record_error("cfg5_full_shape_triples", fwd_ms=t0[0].elapsed_time(t0[1]), bwd_ms=t0[1].elapsed_time(t0[2]))
    assert lg.shape == (1, 50, 10)
    assert torch.isfinite(lg).all() and torch.isfinite(S.grad).all() and torch.isfinite(Q.grad).all()
    assert (lg.argmax(-1) == ep.query_labels).all()
    assert head.transformers[0].k_linear.weight.grad.abs().sum().item() > 0
    assert S.grad.abs().sum().item() > 0 and Q.grad.abs().sum().item() > 0
    with torch.no_grad():
        one = head(ep.support, ep.support_labels, ep.query[:, 7:8].contiguous())["logits"]
    rel = ((one[0, 0] - lg[0, 7].detach()).abs().max() / lg[0, 7].detach().abs().max()).item()
    record_error("cfg5_full_shape_triples", single_query_vs_batch_rel=rel)
    assert rel < 1e-4
