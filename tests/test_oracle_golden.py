"""Pins the oracle (oracle/) against fixtures produced by the reference code itself
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.losses import RECIPE_NAMES, Recipes

G = os.path.join(os.path.dirname(__file__), "golden")
CFG = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1,
           soft_loss_weight_support=1, soft_loss_weight_query=1)


def T(x, grad=False):
    return torch.from_numpy(np.asarray(x)).clone().requires_grad_(grad)


def close(a, b, rtol=2e-5, atol=2e-6):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, np.asarray(b), rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def otam():
    return np.load(os.path.join(G, "otam.npz"))


def test_cos_sim_and_cum_dist_cfg1(otam):
    S, Q = T(otam["cfg1_support"]), T(otam["cfg1_query"])
    sim = oracle.cos_sim(Q.reshape(200, 512), S.reshape(40, 512))
    close(sim, otam["cfg1_sim"])
    d = (1 - sim).reshape(25, 8, 5, 8).permute(0, 2, 1, 3)
    for stable in (False, True):
        close(oracle.otam_cum_dist(d, stable=stable), otam["cfg1_cum_q2s"], rtol=1e-4, atol=1e-5)
        close(oracle.otam_cum_dist(d.transpose(-1, -2), stable=stable), otam["cfg1_cum_s2q"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("shape", ["8x8", "5x7", "1x4", "6x1"])
def test_cum_dist_shapes_and_grad(otam, shape):
    for stable in (False, True):
        d = T(otam[f"cum_{shape}_in"], True)
        c = oracle.otam_cum_dist(d, stable=stable)
        close(c, otam[f"cum_{shape}_out"], rtol=1e-4, atol=1e-5)
        c.sum().backward()
        close(d.grad, otam[f"cum_{shape}_grad"], rtol=1e-3, atol=1e-5)


def test_cfg1_otam_kd_forward_backward(otam):
    S, Q = T(otam["cfg1_support"], True), T(otam["cfg1_query"], True)
    probs = oracle.otam_logits(S, T(otam["cfg1_support_labels"]), Q)
    close(probs, otam["cfg1_probs"], rtol=1e-4, atol=1e-6)
    loss = Recipes(CFG).KD(probs, T(otam["cfg1_teacher_logits"]), T(otam["cfg1_query_labels"]))
    close(loss, otam["cfg1_loss"], rtol=1e-5)
    loss.backward()
    close(S.grad, otam["cfg1_grad_support"], rtol=2e-3, atol=1e-7)
    close(Q.grad, otam["cfg1_grad_query"], rtol=2e-3, atol=1e-7)
    assert (probs.argmax(1).numpy() == otam["cfg1_probs"].argmax(1)).all()


def test_reference_finiteness_is_recorded(otam):
    # the reference's un-stabilised soft-min loses its backward at L >= 12;
    # the stabilised oracle stays finite there
    fin = dict(zip(otam["ref_finite_L"].tolist(), otam["ref_finite_bwd"].tolist()))
    assert fin[8] and fin[10] and not fin[12] and not fin[32]
    rs = np.random.RandomState(0)
    d = T(rs.uniform(0.5, 1.5, (2, 2, 32, 32)).astype(np.float32), True)
    c = oracle.otam_cum_dist_stable(d)
    c.sum().backward()
    assert torch.isfinite(c).all() and torch.isfinite(d.grad).all()


def _head(z, prefix):
    return {k: T(z[f"{prefix}_{k}"]) for k in ("Wk", "bk", "Wv", "bv", "gk", "bek")}


def test_trx_small_cardinalities_and_branch():
    z = np.load(os.path.join(G, "trx_small.npz"))
    S, Q = T(z["small_support"], True), T(z["small_query"], True)
    lab = T(z["small_support_labels"])
    heads = []
    for c in (2, 3):
        h = _head(z, f"small_c{c}")
        for v in h.values():
            v.requires_grad_(True)
        h["card"] = c
        heads.append(h)
        lg = oracle.trx_logits(S, lab, Q, h["Wk"], h["bk"], h["Wv"], h["bv"], h["gk"], h["bek"], c, 5,
                               pe=T(z[f"small_c{c}_pe"]))
        close(lg, z[f"small_logits_c{c}"], rtol=1e-4, atol=1e-4)
    out = oracle.trx_branch_logits(S, lab, Q, heads, 5)
    close(out, z["small_logits_branch"], rtol=1e-4, atol=1e-4)
    (out * T(z["small_upstream"])).sum().backward()
    close(S.grad, z["small_grad_support"], rtol=2e-3, atol=2e-5)
    close(Q.grad, z["small_grad_query"], rtol=2e-3, atol=2e-5)
    for h in heads:
        c = h["card"]
        close(h["Wk"].grad, z[f"small_c{c}_gWk"], rtol=2e-3, atol=2e-5)
        close(h["Wv"].grad, z[f"small_c{c}_gWv"], rtol=2e-3, atol=2e-5)
        close(h["bk"].grad, z[f"small_c{c}_gbk"], rtol=2e-3, atol=2e-5)
        close(h["bv"].grad, z[f"small_c{c}_gbv"], rtol=2e-3, atol=2e-5)
        close(h["gk"].grad, z[f"small_c{c}_ggk"], rtol=2e-3, atol=2e-5)
        close(h["bek"].grad, z[f"small_c{c}_gbek"], rtol=2e-3, atol=2e-5)


def test_student_trx_2fcsup_with_shipped_recipe():
    z = np.load(os.path.join(G, "student_heads.npz"))
    lab = T(z["stu_support_labels"])
    h = _head(z, "stu")
    for v in h.values():
        v.requires_grad_(True)
    S1, S2, Q1, Q2 = (T(z[k], True) for k in ("stu_sup1", "stu_sup2", "stu_qry1", "stu_qry2"))
    args = (h["Wk"], h["bk"], h["Wv"], h["bv"], h["gk"], h["bek"], 2, 5)
    kl = oracle.trx_logits(S1, lab, Q1, *args, pe=T(z["stu_pe"]))
    ce = oracle.trx_logits(S2, lab, Q2, *args, pe=T(z["stu_pe"]))
    sup = oracle.support_dk(S2, 5, 1, 8)
    close(kl, z["stu_logits_kl"], rtol=1e-4, atol=1e-4)
    close(ce, z["stu_logits_ce"], rtol=1e-4, atol=1e-4)
    close(sup, z["stu_logits_sup"], rtol=1e-4, atol=1e-3)
    th = _head(z, "tea")
    tkl = oracle.trx_logits(T(z["tea_sup"]), lab, T(z["tea_qry"]), th["Wk"], th["bk"], th["Wv"], th["bv"],
                            th["gk"], th["bek"], 2, 5, pe=T(z["tea_pe"]))
    tsup = oracle.support_dk(T(z["tea_sup"]), 5, 1, 8)
    close(tkl, z["tea_logits_kl"], rtol=1e-4, atol=1e-4)
    close(tsup, z["tea_logits_sup"], rtol=1e-4, atol=1e-3)
    loss = Recipes(CFG).fc_2_sup_dist({"kl": kl, "ce": ce, "sup": sup}, {"kl": tkl, "sup": tsup},
                                      T(z["stu_query_labels"]))
    close(loss, z["loss"], rtol=1e-4)
    loss.backward()
    for name, ten in (("g_sup1", S1), ("g_sup2", S2), ("g_qry1", Q1), ("g_qry2", Q2)):
        close(ten.grad, z[name], rtol=5e-3, atol=1e-6)
    close(h["Wk"].grad, z["g_Wk"], rtol=5e-3, atol=1e-6)
    close(h["Wv"].grad, z["g_Wv"], rtol=5e-3, atol=1e-6)
    close(h["gk"].grad, z["g_gk"], rtol=5e-3, atol=1e-6)


def test_trx_sup_support_similarity():
    z = np.load(os.path.join(G, "student_heads.npz"))
    h = _head(z, "sup")
    sim, ql = oracle.trx_sup_outputs(T(z["stu_sup1"]), T(z["stu_support_labels"]), T(z["stu_qry1"]),
                                     h["Wk"], h["bk"], h["Wv"], h["bv"], h["gk"], h["bek"], 2, 5,
                                     pe=T(z["sup_pe"]))
    close(sim, z["sup_support_set"], rtol=1e-4, atol=1e-5)
    close(ql, z["sup_query"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", RECIPE_NAMES)
def test_every_distiller_recipe(name):
    z = np.load(os.path.join(G, "losses.npz"))
    lab = T(z["labels20"] if name == "support_sim" else z["labels"])

    def grab(prefix):
        if f"{name}__{prefix}" in z.files:
            return T(z[f"{name}__{prefix}"], prefix == "s")
        keys = [k for k in z.files if k.startswith(f"{name}__{prefix}__")]
        return {k.split("__")[-1]: T(z[k], prefix == "s") for k in keys}

    s, t = grab("s"), grab("t")
    loss = getattr(Recipes(CFG), name)(s, t, lab)
    close(loss.reshape(()), z[f"{name}__loss"], rtol=2e-5, atol=1e-6)
    loss.reshape(()).backward()
    if isinstance(s, dict):
        for k, v in s.items():
            g = v.grad if v.grad is not None else torch.zeros_like(v)
            close(g, z[f"{name}__g__{k}"], rtol=1e-3, atol=1e-7)
    else:
        close(s.grad, z[f"{name}__g"], rtol=1e-3, atol=1e-7)


def test_recipe_list_matches_reference_inventory():
    assert len(RECIPE_NAMES) == 24


def test_e_dist_and_cos_heads():
    z = np.load(os.path.join(G, "edist.npz"))
    for name in ("edist", "cos"):
        S, Q = T(z["support"], True), T(z["query"], True)
        lg = oracle.e_dist_logits(S, T(z["support_labels"]), Q, 5)
        close(lg, z[f"{name}_logits"], rtol=1e-5, atol=1e-5)
        (lg * T(z["upstream"])).sum().backward()
        close(S.grad, z[f"{name}_grad_support"], rtol=1e-3, atol=1e-7)
        close(Q.grad, z[f"{name}_grad_query"], rtol=1e-3, atol=1e-7)


def _feature_head_inputs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)          # module level imports numpy / torch only, never the reference
    return mg.feature_head_inputs()


def test_feature_heads_vs_reference_backbones():
    """adaptive max pool + patch mean + fc1 / fc2 of resnet18_2fc.py, res18_2048 of
    resnet18_student.py (trunk = identity in the fixture)."""
    z = np.load(os.path.join(G, "feature_heads.npz"))
    fm_c, fm_t, W, b, up_c, up_t = _feature_head_inputs()
    C_, T_ = T(fm_c, True), T(fm_t, True)
    Wt, bt = [T(W[h], True) for h in range(2)], [T(b[h], True) for h in range(2)]
    pc, pt = oracle.frame_pool(C_), oracle.frame_pool(T_)
    ctx = oracle.feature_heads(pc, Wt, bt, 8)
    tgt = oracle.feature_heads(pt, Wt, bt, 8)
    loss = 0
    for h in range(2):
        close(ctx[h], z[f"context_features_{h + 1}"], rtol=1e-4, atol=1e-5)
        close(tgt[h], z[f"target_features_{h + 1}"], rtol=1e-4, atol=1e-5)
        loss = loss + (ctx[h] * T(up_c[h])).sum() + (tgt[h] * T(up_t[h])).sum()
    close(ctx[0], z["student_context"], rtol=1e-4, atol=1e-5)
    close(tgt[0], z["student_target"], rtol=1e-4, atol=1e-5)
    loss.backward()
    close(torch.stack([bt[0].grad, bt[1].grad]), z["grad_bias"], rtol=1e-4, atol=1e-4)
    gW = torch.stack([Wt[0].grad, Wt[1].grad])
    close(gW[:, :32], z["grad_weight_rows"], rtol=1e-3, atol=1e-4)
    close(gW.double().pow(2).sum((1, 2)).sqrt(), z["grad_weight_norm"], rtol=1e-5, atol=0)
    close(C_.grad[:2], z["grad_fmap_context_head"], rtol=1e-3, atol=1e-5)
    close(T_.grad[:2], z["grad_fmap_target_head"], rtol=1e-3, atol=1e-5)
    close(C_.grad.sum((2, 3)), z["grad_fmap_context_sum"], rtol=1e-3, atol=1e-4)
    close(T_.grad.sum((2, 3)), z["grad_fmap_target_sum"], rtol=1e-3, atol=1e-4)


def test_frame_pool_windows_match_torch_adaptive_pooling():
    for (H, W, o) in ((7, 7, 4), (8, 6, 3), (5, 5, 5), (14, 14, 4)):
        x = torch.randn(3, 4, H, W, generator=torch.Generator().manual_seed(H * 100 + W))
        ref = torch.nn.functional.adaptive_max_pool2d(x, (o, o)).reshape(3, 4, -1).mean(-1)
        close(oracle.frame_pool(x, o), ref, rtol=1e-6, atol=1e-6)


def test_strm_distance_loss_matches_reference():
    """oracle.strm_distance_logits against the reference's own DistanceLoss (strm_res18_sup.py:162-243), forward,
    every gradient, and the ragged / missing-class case."""
    import oracle
    z = np.load(os.path.join(G, "strm.npz"))
    T = lambda k, g=False: torch.from_numpy(z[k]).clone().requires_grad_(g)
    S, Q, W, b = T("support", True), T("query", True), T("W", True), T("b", True)
    lg = oracle.strm_distance_logits(S, T("support_labels"), Q, W, b, 2, 5)
    np.testing.assert_allclose(lg.detach().numpy(), z["logits"], rtol=1e-5, atol=1e-5)
    (lg * T("upstream")).sum().backward()
    for name, x in (("grad_support", S), ("grad_query", Q), ("gW", W), ("gb", b)):
        np.testing.assert_allclose(x.grad.numpy(), z[name], rtol=1e-4, atol=1e-5 * np.abs(z[name]).max())
    keep = z["ragged_keep"]
    lg2 = oracle.strm_distance_logits(T("support")[keep], T("support_labels")[keep], T("query"), T("W"), T("b"), 2, 5)
    np.testing.assert_allclose(lg2.numpy(), z["ragged_logits"], rtol=1e-5, atol=1e-5)
    assert (z["ragged_logits"][:, 3] == 0).all()          # the absent class keeps the zero of dist_all (:212)
