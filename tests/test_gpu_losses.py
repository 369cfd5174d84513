"""GPU parity of the D2M losses through the C-ABI: every Distiller recipe against the reference's
own outputs (tests/golden/losses.npz), the fused feature-MSE pass, accuracy, batched episodes.
All loss arithmetic is fp32: values rel 1e-4 (north_star bound: 1e-3), gradients rel-L2 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle.losses import RECIPE_NAMES

from conftest import allclose, assert_close

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


@pytest.mark.parametrize("name", RECIPE_NAMES)
def test_every_recipe_vs_reference(name):
    import distillers
    d = dev()
    z = np.load(os.path.join(G, "losses.npz"))
    lab = T(z["labels20"] if name == "support_sim" else z["labels"], device=d)

    def grab(prefix):
        if f"{name}__{prefix}" in z.files:
            return T(z[f"{name}__{prefix}"], prefix == "s", d)
        keys = [k for k in z.files if k.startswith(f"{name}__{prefix}__")]
        return {k.split("__")[-1]: T(z[k], prefix == "s", d) for k in keys}

    s, t = grab("s"), grab("t")
    res = getattr(distillers.Distiller(name, dict(CFG), d), name)(s, t, lab)
    ref = float(z[f"{name}__loss"])
    assert abs(res["loss"].item() - ref) <= 1e-4 * max(abs(ref), 1e-3), (res["loss"].item(), ref)
    res["loss"].backward()
    items = s.items() if isinstance(s, dict) else [(None, s)]
    for k, v in items:
        g_ref = z[f"{name}__g__{k}"] if k is not None else z[f"{name}__g"]
        g = v.grad.cpu().numpy() if v.grad is not None else np.zeros_like(g_ref)
        denom = max(np.linalg.norm(g_ref), 1e-12)
        assert np.linalg.norm(g - g_ref) / denom < 1e-4 or np.abs(g - g_ref).max() < 1e-7, (name, k)
    for k, v in res.items():       # the reference's reported parts, where it reports them
        key = f"{name}__part__{k}"
        if key in z.files and torch.is_tensor(v) and z[key].size == 1:
            assert abs(v.item() - float(z[key].reshape(-1)[0])) <= 1e-4 * max(abs(float(z[key].reshape(-1)[0])), 1e-3), key


def test_module_level_functions_vs_oracle():
    import oracle
    import distillers
    d = dev()
    rs = np.random.RandomState(0)
    s, t = rs.standard_normal((25, 5)).astype(np.float32) * 3, rs.standard_normal((25, 5)).astype(np.float32) * 3
    a = distillers.kd_loss(T(s, device=d), T(t, device=d), 4).item()
    b = oracle.kd_loss(T(s), T(t), 4).item()
    assert abs(a - b) <= 1e-5 * abs(b)
    a = distillers.inter_class_relation(T(s, device=d), T(t, device=d)).item()
    b = oracle.inter_class_relation(T(s), T(t)).item()
    assert abs(a - b) <= 1e-5 * abs(b)


def test_batched_episodes_sum_like_gradient_accumulation():
    """[B, rows, cols] logits: loss = sum over episodes, each episode with its own focal weight."""
    import distillers
    from oracle.losses import Recipes
    d = dev()
    rs = np.random.RandomState(1)
    B = 7
    mk = lambda *sh: (3 * rs.standard_normal(sh)).astype(np.float32)
    s_np = {"kl": mk(B, 25, 5), "ce": mk(B, 25, 5), "sup": 40 * mk(B, 5, 4)}
    t_np = {"kl": mk(B, 25, 5), "sup": 40 * mk(B, 5, 4)}
    y = rs.randint(0, 5, (B, 25)).astype(np.int64)
    for name in ("fc_2_sup_dist", "fc_2_sup_dist_wsl", "fc_2_sup"):
        s = {k: T(v, True, d) for k, v in s_np.items()}
        t = {k: T(v, device=d) for k, v in t_np.items()}
        res = getattr(distillers.Distiller(name, dict(CFG), d), name)(s, t, T(y, device=d))
        res["loss"].backward()
        tot, grads = 0.0, {k: [] for k in s_np}
        for b in range(B):
            sb = {k: T(v[b], True) for k, v in s_np.items()}
            lb = getattr(Recipes(CFG), name)(sb, {k: T(v[b]) for k, v in t_np.items()}, T(y[b]))
            lb.backward()
            tot += lb.item()
            for k in sb:
                grads[k].append(sb[k].grad)
        assert abs(res["loss"].item() - tot) <= 1e-4 * abs(tot)
        for k in s:
            ref = torch.stack(grads[k])
            assert ((s[k].grad.cpu() - ref).norm() / ref.norm()).item() < 1e-4


@pytest.mark.parametrize("n,dtype", [(10 * 8 * 64, torch.float32), (50 * 8 * 2048 * 3 + 3, torch.float32),
                                     (50 * 8 * 2048 * 2, torch.bfloat16), (1003, torch.bfloat16)])
def test_feature_mse_single_pass(n, dtype):
    from lmkd import ops
    d = dev()
    g = torch.Generator().manual_seed(n)
    s = torch.randn(n, generator=g).to(dtype)
    t = torch.randn(n, generator=g).to(dtype)
    sg = s.to(d).requires_grad_(True)
    loss = ops.feature_mse(sg, t.to(d), weight=1.5)
    (loss * 1.0).backward()
    ref_s = s.double().requires_grad_(True)
    ref = 1.5 * ((ref_s - t.double()) ** 2).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * ref.item()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert ((sg.grad.double().cpu() - ref_s.grad).norm() / ref_s.grad.norm()).item() < tol
    # a non-unit upstream gradient goes through the device-scalar scale kernel
    sg2 = s.to(d).requires_grad_(True)
    (ops.feature_mse(sg2, t.to(d), weight=1.5) * 0.25).backward()
    assert ((sg2.grad.double().cpu() - 0.25 * ref_s.grad).norm() / ref_s.grad.norm()).item() < tol


def test_kl_feature_recipe_batched_cfg3_shape():
    """BASELINE config 3 shape at reduced batch: features [B, 50, 8, 2048], summed teacher streams."""
    import distillers
    from oracle.losses import Recipes
    from lmkd.episodes import make_episodes
    d = dev()
    B = 3
    ep = make_episodes(B, 5, 5, 5, 8, 2048, modalities=3, seed=5)
    sf = torch.cat([ep.support, ep.query], 1)
    tf = torch.cat([ep.teacher_support, ep.teacher_query], 1)
    rs = np.random.RandomState(3)
    lg, tl = T(3 * rs.standard_normal((B, 25, 5)).astype(np.float32)), T(3 * rs.standard_normal((B, 25, 5)).astype(np.float32))
    sfg, lgg = sf.to(d).requires_grad_(True), lg.to(d).requires_grad_(True)
    res = distillers.Distiller("KL_feature", dict(CFG), d).KL_feature(
        {"logits": lgg, "feature": sfg}, {"logits": tl.to(d), "feature": tf.to(d)}, ep.query_labels.to(d))
    res["loss"].backward()
    tot = 0.0
    gref = []
    for b in range(B):
        s_b, l_b = sf[b].clone().requires_grad_(True), lg[b].clone().requires_grad_(True)
        lb = Recipes(CFG).KL_feature({"logits": l_b, "feature": s_b}, {"logits": tl[b], "feature": tf[b]},
                                     ep.query_labels[b])
        lb.backward()
        tot += lb.item()
        gref.append(s_b.grad)
    assert abs(res["loss"].item() - tot) <= 1e-4 * abs(tot)
    gref = torch.stack(gref)
    assert ((sfg.grad.cpu() - gref).norm() / gref.norm()).item() < 1e-5


def test_accuracy_is_bit_exact():
    import utils
    d = dev()
    rs = np.random.RandomState(2)
    lg = rs.standard_normal((4096, 5)).astype(np.float32)
    y = rs.randint(0, 5, 4096).astype(np.int64)
    acc = utils.aggregate_accuracy(T(lg, device=d), T(y, device=d)).item()
    ref = float((lg.argmax(1) == y).mean())
    assert acc == pytest.approx(ref, abs=0) or abs(acc - ref) < 1e-7
    assert int(round(acc * 4096)) == int((lg.argmax(1) == y).sum())


def test_support_dk_vs_oracle():
    import oracle
    from lmkd import ops
    d = dev()
    rs = np.random.RandomState(4)
    x = rs.standard_normal((2, 15, 8, 256)).astype(np.float32)
    up = rs.standard_normal((2, 5, 4)).astype(np.float32)
    xg = T(x, True, d)
    out = ops.support_dk(xg, 5, 3)
    (out * T(up, device=d)).sum().backward()
    for b in range(2):
        xb = T(x[b], True)
        ref = oracle.support_dk(xb, 5, 3, 8)
        (ref * T(up[b])).sum().backward()
        assert_close(out[b].detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-4)
        assert ((xg.grad[b].cpu() - xb.grad).norm() / xb.grad.norm()).item() < 1e-5


def test_e_dist_and_cos_heads_vs_reference():
    """Frame-mean Euclidean heads (e_dist.py / COS.py) against the reference's outputs and gradients."""
    import types
    import model.classifiers as C
    d = dev()
    z = np.load(os.path.join(G, "edist.npz"))
    args = types.SimpleNamespace(seq_len=8, way=5, shot=2)
    for name, cls in (("edist", C.e_dist), ("cos", C.CosDistance)):
        S, Q = T(z["support"], True, d), T(z["query"], True, d)
        o = cls(args)(S, T(z["support_labels"], device=d), Q)
        lg = o["logits"] if isinstance(o, dict) else o
        assert_close(lg.detach().cpu().numpy(), z[f"{name}_logits"], rtol=1e-5, atol=1e-4)
        (lg * T(z["upstream"], device=d)).sum().backward()
        for g, ref in ((S.grad, z[f"{name}_grad_support"]), (Q.grad, z[f"{name}_grad_query"])):
            assert np.linalg.norm(g.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-5
    # batched episodes + the two-head wrapper
    from lmkd.episodes import make_episodes
    import oracle
    ep = make_episodes(3, 5, 5, 5, 8, 512, teacher_dim=8, seed=12)
    args = types.SimpleNamespace(seq_len=8, way=5, shot=5)
    out = C.e_dist_1fc_sup(args)(ep.support.to(d), ep.support_labels.to(d), ep.query.to(d))["logits"]
    for b in range(3):
        ref = oracle.e_dist_logits(ep.support[b], ep.support_labels[b], ep.query[b], 5)
        assert_close(out["kl"][b].cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-4)
    assert out["sup"].shape == (3, 5, 4)


def test_feature_mse_refuses_a_second_backward():
    """The fused kernel hands out its gradient buffer once (the upstream scalar is applied in place): a second
    backward through the same graph must raise instead of returning g^2 * ds."""
    from lmkd import ops
    d = dev()
    s = torch.randn(2, 10, 8, 64, device=d, requires_grad=True)
    t = torch.randn(2, 10, 8, 64, device=d)
    loss = ops.feature_mse(s, t) * 3.0
    loss.backward(retain_graph=True)
    ref = 3.0 * 2.0 * (s.detach() - t) / s.numel()
    assert allclose(s.grad, ref, rtol=1e-5, atol=1e-7)
    with pytest.raises(RuntimeError, match="twice"):
        loss.backward()
