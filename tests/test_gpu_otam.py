"""GPU parity of the OTAM path (similarity GEMM, wavefront DP, class softmax, backward) through
the C-ABI, against the golden fixtures made from the reference and against the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import allclose, assert_close, rel_l2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
CFG = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1,
           soft_loss_weight_support=1, soft_loss_weight_query=1)


def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def T(x, grad=False, device=None):
    t = torch.from_numpy(np.asarray(x)).clone()
    if device is not None:
        t = t.to(device)
    return t.requires_grad_(grad)


@pytest.fixture(scope="module")
def otam():
    return np.load(os.path.join(G, "otam.npz"))


@pytest.mark.parametrize("shape", ["8x8", "5x7", "1x4", "6x1"])
def test_cum_dist_matches_reference(otam, shape):
    """fp32 recurrence: the stabilised soft-min equals the reference's wherever that is finite."""
    from lmkd import ops
    d = T(otam[f"cum_{shape}_in"], device=dev())
    out, gd = ops.otam_cum_dist(d, 0.1, grad_out=torch.ones(d.shape[:2], device=dev()))
    assert_close(out.cpu().numpy(), otam[f"cum_{shape}_out"], rtol=1e-4, atol=1e-5)
    assert_close(gd.cpu().numpy(), otam[f"cum_{shape}_grad"], rtol=2e-3, atol=1e-5)


def test_cum_dist_long_clip_is_finite_and_matches_oracle():
    """L = 32: the reference's backward is NaN here; compare with the stabilised fp64 oracle."""
    import oracle
    from lmkd import ops
    rs = np.random.RandomState(5)
    d_np = rs.uniform(0.2, 1.8, (6, 3, 32, 32)).astype(np.float32)
    d64 = torch.from_numpy(d_np).double().requires_grad_(True)
    ref = oracle.otam_cum_dist_stable(d64)
    ref.sum().backward()
    out, gd = ops.otam_cum_dist(T(d_np, device=dev()), 0.1, grad_out=torch.ones(6, 3, device=dev()))
    assert torch.isfinite(out).all() and torch.isfinite(gd).all()
    assert_close(out.cpu().numpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-4)
    assert rel_l2(gd, d64.grad) < 1e-3


def test_frame_dists_cfg1(otam):
    """bf16 tcgen05 contraction with fp32 norms: |dist - ref| small on cosine scale."""
    from lmkd import ops
    q = T(otam["cfg1_query"], device=dev()).reshape(1, 200, 512)
    s = T(otam["cfg1_support"], device=dev()).reshape(1, 40, 512)
    dist = ops.frame_dists(q, s)
    ref = 1.0 - otam["cfg1_sim"]
    assert np.abs(dist[0].cpu().numpy() - ref).max() < 2e-3


def test_cfg1_otam_kd_forward_backward(otam):
    """BASELINE config 1: 5-way 1-shot, 8 frames x 512-d, OTAM head + Distiller.KD, fwd+bwd.
    Tolerances: probabilities rel 1e-2, loss rel 1e-3 (north_star), gradients rel-L2 1e-2,
    argmax bit-exact."""
    import distillers
    import model.classifiers as C
    import types
    d = dev()
    args = types.SimpleNamespace(seq_len=8, way=5, shot=1)
    S, Q = T(otam["cfg1_support"], True, d), T(otam["cfg1_query"], True, d)
    S.retain_grad(), Q.retain_grad()
    head = C.OTAM(args)
    probs = head(S, T(otam["cfg1_support_labels"], device=d), Q)["logits"]
    assert_close(probs.detach().cpu().numpy(), otam["cfg1_probs"], rtol=1e-2, atol=1e-5)
    assert (probs.argmax(1).cpu().numpy() == otam["cfg1_probs"].argmax(1)).all()
    loss = distillers.Distiller("KD", CFG, d).KD(probs, T(otam["cfg1_teacher_logits"], device=d),
                                                 T(otam["cfg1_query_labels"], device=d))["loss"]
    assert abs(loss.item() - float(otam["cfg1_loss"])) <= 1e-3 * abs(float(otam["cfg1_loss"]))
    loss.backward()
    assert rel_l2(S.grad, otam["cfg1_grad_support"]) < 1e-2
    assert rel_l2(Q.grad, otam["cfg1_grad_query"]) < 1e-2


@pytest.mark.parametrize("B,way,shot,qpc,L,D", [(3, 5, 5, 5, 8, 2048), (2, 10, 5, 2, 32, 256), (2, 3, 2, 1, 5, 64),
                                               (1, 4, 3, 2, 16, 128)])
def test_batched_otam_vs_oracle(B, way, shot, qpc, L, D):
    """Batched episodes incl. BASELINE config 4 / 5 shapes; oracle = stabilised fp32 restatement."""
    import oracle
    from lmkd import ops
    from lmkd.episodes import make_episodes
    d = dev()
    ep = make_episodes(B, way, shot, qpc, L, D, teacher_dim=D, seed=11)
    up = torch.randn(B, way * qpc, way, generator=torch.Generator().manual_seed(3))
    S, Q = ep.support.to(d).requires_grad_(True), ep.query.to(d).requires_grad_(True)
    probs = ops.otam_probs(S, ep.support_labels.to(d), Q, way)
    (probs * up.to(d)).sum().backward()
    for b in range(B):
        s, q = ep.support[b].clone().requires_grad_(True), ep.query[b].clone().requires_grad_(True)
        ref = oracle.otam_logits(s, ep.support_labels[b], q, stable=True)
        (ref * up[b]).sum().backward()
        assert_close(probs[b].detach().cpu().numpy(), ref.detach().numpy(), rtol=2e-3, atol=1e-4)
        assert (probs[b].argmax(1).cpu() == ref.argmax(1)).all()
        assert rel_l2(S.grad[b], s.grad) < 1e-2
        assert rel_l2(Q.grad[b], q.grad) < 1e-2


def test_nan_guard_returns_zero_logits():
    """CNN_OTAM returns all-zero logits when a support feature is NaN (model.py:3322-3324)."""
    from lmkd import ops
    from lmkd.episodes import make_episodes
    d = dev()
    ep = make_episodes(2, 5, 1, 1, 8, 64, teacher_dim=64, seed=2)
    sup = ep.support.clone()
    sup[1, 2, 3, 4] = float("nan")
    probs = ops.otam_probs(sup.to(d), ep.support_labels.to(d), ep.query.to(d), 5)
    assert torch.isfinite(probs[0]).all() and abs(probs[0].sum().item() - 5.0) < 1e-3
    assert (probs[1] == 0).all()


def test_nan_guard_backward_gives_zero_feature_gradients():
    """The reference returns a DETACHED zero tensor for a NaN episode (model.py:3322-3324), so no gradient reaches
    its features; the other episodes of the batch are unaffected."""
    from lmkd import ops
    from lmkd.episodes import make_episodes
    d = dev()
    ep = make_episodes(3, 5, 1, 2, 8, 64, teacher_dim=64, seed=2)
    sup = ep.support.clone()
    sup[1, 2, 3, 4] = float("nan")
    S, Q = sup.to(d).requires_grad_(True), ep.query.to(d).requires_grad_(True)
    probs = ops.otam_probs(S, ep.support_labels.to(d), Q, 5)
    up = torch.randn(probs.shape, generator=torch.Generator().manual_seed(0)).to(d)
    (probs * up).sum().backward()
    assert torch.isfinite(S.grad).all() and torch.isfinite(Q.grad).all()
    assert (S.grad[1] == 0).all() and (Q.grad[1] == 0).all()
    assert S.grad[0].abs().sum().item() > 0 and Q.grad[2].abs().sum().item() > 0
    # the clean episodes match a run without the poisoned one
    S2 = ep.support[[0, 2]].to(d).requires_grad_(True)
    Q2 = ep.query[[0, 2]].to(d).requires_grad_(True)
    p2 = ops.otam_probs(S2, ep.support_labels[[0, 2]].to(d), Q2, 5)
    (p2 * up[[0, 2]]).sum().backward()
    assert allclose(S.grad[[0, 2]], S2.grad, rtol=1e-5, atol=1e-6)


def test_bad_label_raises_on_the_next_call_without_an_explicit_check():
    """The status word lives in pinned host memory: the wrapper that FOLLOWS the offending launch sees it."""
    from lmkd import ops
    from lmkd.episodes import make_episodes
    d = dev()
    ep = make_episodes(1, 5, 1, 1, 8, 64, teacher_dim=64, seed=2)
    lab = ep.support_labels.clone()
    lab[0, 0] = 9.0
    ops.otam_probs(ep.support.to(d), lab.to(d), ep.query.to(d), 5)
    torch.cuda.synchronize()        # only so that "the kernel has run" is deterministic in the test
    with pytest.raises(RuntimeError, match="outside"):
        ops.otam_probs(ep.support.to(d), ep.support_labels.to(d), ep.query.to(d), 5)
    ops.otam_probs(ep.support.to(d), ep.support_labels.to(d), ep.query.to(d), 5)      # reported once, then clear


def test_out_of_range_label_is_reported():
    from lmkd import check_device_status, ops
    from lmkd.episodes import make_episodes
    d = dev()
    ep = make_episodes(1, 5, 1, 1, 8, 64, teacher_dim=64, seed=2)
    lab = ep.support_labels.clone()
    lab[0, 0] = 9.0
    ops.otam_probs(ep.support.to(d), lab.to(d), ep.query.to(d), 5)
    with pytest.raises(RuntimeError):
        check_device_status(d)
    check_device_status(d)   # cleared


def test_cpu_tensor_is_rejected():
    from lmkd import ops
    with pytest.raises(RuntimeError):
        ops.otam_probs(torch.zeros(1, 5, 8, 64), torch.zeros(1, 5), torch.zeros(1, 5, 8, 64), 5)
