"""Generate the golden fixtures in this directory by running the REFERENCE code itself.

Run in the authoring container only (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py [/root/reference]

It imports the reference's own modules (student `model.classifiers`, `distillers`, and the
teacher-side `teacher/code/model.py` OTAM / multi-cardinality TRX) on CPU under the two import
shims described in SURVEY.md §8c, feeds them seeded inputs (numpy RandomState, stream-stable
across versions) and stores inputs + outputs + autograd gradients as small .npz files.
The tests never import the reference; they only read these files.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 3483   # the reference's own constant (model/classifiers/TRX.py:18)


def load_reference(root: str):
    sys.path.insert(0, root)
    torch.Tensor.cuda = lambda self, *a, **k: self          # shim 1: TRX.py:72 calls .cuda()
    import distillers                                        # noqa
    import model.classifiers as C                            # noqa
    for n in ("timm", "turtle", "matplotlib", "matplotlib.style"):   # shim 2
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["turtle"].forward = None
    sys.modules["matplotlib.style"].context = None

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    tdir = os.path.join(root, "teacher", "code")
    sys.path.insert(0, tdir)
    sys.modules.pop("utils", None)
    load("utils", os.path.join(tdir, "utils.py"))
    load("transformer", os.path.join(tdir, "transformer.py"))
    T = load("teacher_model", os.path.join(tdir, "model.py"))
    return distillers, C, T


def structured_episode(rs, way, shot, nq_per_class, L, D, noise=0.5, shuffle=True):
    """Class-structured synthetic episode: centroid + noise, shuffled order."""
    cent = rs.standard_normal((way, L, D)).astype(np.float32)
    s_lab = np.repeat(np.arange(way), shot)
    q_lab = np.repeat(np.arange(way), nq_per_class)
    if shuffle:
        s_lab = s_lab[rs.permutation(len(s_lab))]
        q_lab = q_lab[rs.permutation(len(q_lab))]
    sup = cent[s_lab] + noise * rs.standard_normal((len(s_lab), L, D)).astype(np.float32)
    qry = cent[q_lab] + noise * rs.standard_normal((len(q_lab), L, D)).astype(np.float32)
    return sup.astype(np.float32), s_lab.astype(np.float32), qry.astype(np.float32), q_lab.astype(np.int64)


def t(x, grad=False):
    return torch.from_numpy(np.ascontiguousarray(x)).clone().requires_grad_(grad)


def npy(x):
    return x.detach().cpu().numpy()


def gen_otam(T, distillers):
    out = {}
    rs = np.random.RandomState(SEED)
    # --- cfg1: 5-way 1-shot, 25 queries, L=8, D=512, OTAM + Distiller.KD ----------------
    sup, s_lab, qry, q_lab = structured_episode(rs, 5, 1, 5, 8, 512)
    teacher_logits = rs.standard_normal((25, 5)).astype(np.float32)
    S, Q = t(sup, True), t(qry, True)
    head = T.CNN_OTAM()
    probs = head(S, t(s_lab), Q)["logits"]
    cfg = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1,
               soft_loss_weight_support=1, soft_loss_weight_query=1)
    dist = distillers.Distiller("KD", cfg, "cpu")
    loss = dist.KD(probs, t(teacher_logits), t(q_lab))["loss"]
    loss.backward()
    with torch.no_grad():
        sim = T.cos_sim(Q.reshape(200, 512), S.reshape(40, 512))
        d = (1 - sim).reshape(25, 8, 5, 8).permute(0, 2, 1, 3)
        cum_a = T.OTAM_cum_dist(d)
        cum_b = T.OTAM_cum_dist(d.transpose(-1, -2))
    out.update(cfg1_support=sup, cfg1_support_labels=s_lab, cfg1_query=qry, cfg1_query_labels=q_lab,
               cfg1_teacher_logits=teacher_logits, cfg1_sim=npy(sim), cfg1_cum_q2s=npy(cum_a),
               cfg1_cum_s2q=npy(cum_b), cfg1_probs=npy(probs), cfg1_loss=npy(loss),
               cfg1_grad_support=npy(S.grad), cfg1_grad_query=npy(Q.grad))
    # --- raw OTAM_cum_dist on random [3,4,L,M] incl. non-square and its gradient --------
    for (L, M) in [(8, 8), (5, 7), (1, 4), (6, 1)]:
        d = t(rs.uniform(0, 2, (3, 4, L, M)).astype(np.float32), True)
        c = T.OTAM_cum_dist(d)
        c.sum().backward()
        out[f"cum_{L}x{M}_in"] = npy(d)
        out[f"cum_{L}x{M}_out"] = npy(c)
        out[f"cum_{L}x{M}_grad"] = npy(d.grad)
    # --- where the reference stops being finite (documented in DESIGN.md) --------------
    fin = {}
    for L in (8, 10, 12, 16, 32):
        sup, s_lab, qry, q_lab = structured_episode(np.random.RandomState(SEED + L), 5, 1, 1, L, 64)
        S, Q = t(sup, True), t(qry, True)
        p = T.CNN_OTAM()(S, t(s_lab), Q)["logits"]
        p.sum().backward() if torch.isfinite(p).all() else None
        fin[L] = (bool(torch.isfinite(p).all()), bool(S.grad is not None and torch.isfinite(S.grad).all()))
    out["ref_finite_L"] = np.array(sorted(fin), dtype=np.int64)
    out["ref_finite_fwd"] = np.array([fin[k][0] for k in sorted(fin)])
    out["ref_finite_bwd"] = np.array([fin[k][1] for k in sorted(fin)])
    np.savez_compressed(os.path.join(HERE, "otam.npz"), **out)
    print("otam.npz", {k: v.shape for k, v in out.items() if k.startswith("cfg1")}, fin)


def head_state(m):
    return dict(Wk=npy(m.k_linear.weight), bk=npy(m.k_linear.bias), Wv=npy(m.v_linear.weight),
                bv=npy(m.v_linear.bias), gk=npy(m.norm_k.weight), bek=npy(m.norm_k.bias),
                pe=npy(m.pe.pe[0]))


def randomize_head(m, rs):
    """Non-trivial LayerNorm affine + reference-scale Linear init, from the stable stream."""
    with torch.no_grad():
        for p in (m.k_linear.weight, m.v_linear.weight):
            bound = 1.0 / np.sqrt(p.shape[1])
            p.copy_(t(rs.uniform(-bound, bound, tuple(p.shape)).astype(np.float32)))
        for p in (m.k_linear.bias, m.v_linear.bias, m.norm_k.bias):
            p.copy_(t(rs.uniform(-0.1, 0.1, tuple(p.shape)).astype(np.float32)))
        m.norm_k.weight.copy_(t(rs.uniform(0.5, 1.5, tuple(m.norm_k.weight.shape)).astype(np.float32)))


def gen_trx(T, C):
    out = {}
    rs = np.random.RandomState(SEED + 1)
    # --- teacher-side generic TemporalCrossTransformer, small dims, c = 2 and 3, TrxBranch mean
    L, D, d = 8, 64, 32
    args = types.SimpleNamespace(seq_len=L, trans_dropout=0.1, trans_linear_out_dim=d,
                                 trans_linear_in_dim=D, way=5, shot=3, temp_set=[2, 3], num_gpus=1)
    sup, s_lab, qry, q_lab = structured_episode(rs, 5, 3, 2, L, D)
    branch = T.TrxBranch(args).eval()
    for i, m in enumerate(branch.transformers):
        randomize_head(m, rs)
        for k, v in head_state(m).items():
            out[f"small_c{m.temporal_set_size}_{k}"] = v
    S, Q = t(sup, True), t(qry, True)
    per_card = [m(S, t(s_lab), Q)["logits"] for m in branch.transformers]
    logits = branch(S, Q, t(s_lab))["logits"][0]
    upstream = rs.standard_normal((10, 5)).astype(np.float32)
    (logits * t(upstream)).sum().backward()
    out.update(small_upstream=upstream, small_support=sup, small_support_labels=s_lab, small_query=qry, small_query_labels=q_lab,
               small_logits_c2=npy(per_card[0]), small_logits_c3=npy(per_card[1]),
               small_logits_branch=npy(logits), small_grad_support=npy(S.grad), small_grad_query=npy(Q.grad))
    for m in branch.transformers:
        c = m.temporal_set_size
        out[f"small_c{c}_gWk"] = npy(m.k_linear.weight.grad)
        out[f"small_c{c}_gbk"] = npy(m.k_linear.bias.grad)
        out[f"small_c{c}_gWv"] = npy(m.v_linear.weight.grad)
        out[f"small_c{c}_gbv"] = npy(m.v_linear.bias.grad)
        out[f"small_c{c}_ggk"] = npy(m.norm_k.weight.grad)
        out[f"small_c{c}_gbek"] = npy(m.norm_k.bias.grad)
    np.savez_compressed(os.path.join(HERE, "trx_small.npz"), **out)
    print("trx_small.npz", out["small_logits_branch"][:2])


def gen_student(C):
    """Student-side heads (D fixed at 2048 by TRX.py:59-62) with a small key dim, 5-way 1-shot."""
    out = {}
    rs = np.random.RandomState(SEED + 2)
    L, d = 8, 16
    args = types.SimpleNamespace(seq_len=L, trans_dropout=0.1, trans_linear_out_dim=d,
                                 trans_linear_in_dim=2048, way=5, shot=1, temp_set=[2], num_gpus=1,
                                 device="cpu")
    sup1, s_lab, qry1, q_lab = structured_episode(rs, 5, 1, 1, L, 2048, shuffle=False)
    sup2 = sup1 + 0.1 * rs.standard_normal(sup1.shape).astype(np.float32)
    qry2 = qry1 + 0.1 * rs.standard_normal(qry1.shape).astype(np.float32)
    head = C.TRX_2fcsup(args).eval()
    randomize_head(head.transformers, rs)
    for k, v in head_state(head.transformers).items():
        out[f"stu_{k}"] = v
    S1, S2, Q1, Q2 = t(sup1, True), t(sup2, True), t(qry1, True), t(qry2, True)
    lg = head({"context_features_1": S1, "context_features_2": S2}, t(s_lab),
              {"target_features_1": Q1, "target_features_2": Q2})["logits"]
    teacher = C.TRX_2fcsup_fixed(args).eval()
    randomize_head(teacher.transformers, rs)
    for k, v in head_state(teacher.transformers).items():
        out[f"tea_{k}"] = v
    tsup = sup1 + 0.3 * rs.standard_normal(sup1.shape).astype(np.float32)
    tqry = qry1 + 0.3 * rs.standard_normal(qry1.shape).astype(np.float32)
    tl = teacher(t(tsup), t(s_lab), t(tqry))["logits"]
    import distillers
    cfg = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1,
               soft_loss_weight_support=1, soft_loss_weight_query=1)
    loss = distillers.Distiller("fc_2_sup_dist", cfg, "cpu").fc_2_sup_dist(lg, tl, t(q_lab))
    loss["loss"].backward()
    out.update(stu_sup1=sup1, stu_sup2=sup2, stu_qry1=qry1, stu_qry2=qry2, stu_support_labels=s_lab,
               stu_query_labels=q_lab, tea_sup=tsup, tea_qry=tqry,
               stu_logits_kl=npy(lg["kl"]), stu_logits_ce=npy(lg["ce"]), stu_logits_sup=npy(lg["sup"]),
               tea_logits_kl=npy(tl["kl"]), tea_logits_sup=npy(tl["sup"]),
               loss=npy(loss["loss"]), soft_loss=npy(loss["soft_loss"]), hard_loss=npy(loss["hard_loss"]),
               g_sup1=npy(S1.grad), g_sup2=npy(S2.grad), g_qry1=npy(Q1.grad), g_qry2=npy(Q2.grad),
               g_Wk=npy(head.transformers.k_linear.weight.grad), g_Wv=npy(head.transformers.v_linear.weight.grad),
               g_gk=npy(head.transformers.norm_k.weight.grad))
    # TRX_sup: per-class prototype cosine matrix + query logits (TRX_sup.py:114-179)
    hs = C.TRX_sup(args).eval()
    randomize_head(hs.transformers, rs)
    for k, v in head_state(hs.transformers).items():
        out[f"sup_{k}"] = v
    o = hs(t(sup1), t(s_lab), t(qry1))["logits"]
    out.update(sup_support_set=npy(o["support_set"]), sup_query=npy(o["query"]))
    np.savez_compressed(os.path.join(HERE, "student_heads.npz"), **out)
    print("student_heads.npz loss", out["loss"], "kl logits", out["stu_logits_kl"][0])


def gen_losses(distillers):
    """Every Distiller recipe on random logits; values + gradients w.r.t. all student inputs."""
    out = {}
    rs = np.random.RandomState(SEED + 3)
    cfg = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1,
               soft_loss_weight_support=1, soft_loss_weight_query=1)
    nq, w = 25, 5
    y = rs.randint(0, w, nq).astype(np.int64)

    def L(scale=3.0, shape=(nq, w)):
        return (scale * rs.standard_normal(shape)).astype(np.float32)

    keysets = {
        "tensor": (lambda: L(), lambda: L()),
        "fc": (lambda: {"fc_1": L(), "fc_2": L()}, lambda: L()),
        "strm": (lambda: {"pat": L(), "fr": L()}, lambda: L()),
        "klcesup": (lambda: {"kl": L(), "ce": L(), "sup": L(40.0, (5, 4))},
                    lambda: {"kl": L(), "sup": L(40.0, (5, 4))}),
        "sup2": (lambda: {"kl": L(), "ce": L(), "sup_kl": L(40.0, (5, 4)), "sup_ce": L(40.0, (5, 4))},
                 lambda: {"kl": L(), "sup": L(40.0, (5, 4))}),
        "strmsup": (lambda: {"pat": L(), "fr": L(), "fr1": L(), "fr2": L(), "sup": L(40.0, (5, 4))},
                    lambda: {"kl": L(), "sup": L(40.0, (5, 4))}),
        "klsup": (lambda: {"kl": L(), "sup": L(40.0, (5, 4))}, lambda: {"kl": L(), "sup": L(40.0, (5, 4))}),
        "supsim": (lambda: {"support_set": L(1.0, (20, 5, 5)), "query": L(3.0, (20, 5))},
                   lambda: {"support_set": L(1.0, (20, 5, 5)), "query": L(3.0, (20, 5))}),
        "feat": (lambda: {"logits": L(), "feature": L(1.0, (10, 8, 64))},
                 lambda: {"logits": L(), "feature": L(1.0, (10, 8, 64))}),
    }
    recipe_keys = dict(KD="tensor", wsl="tensor", ce="tensor", Dist_KD="tensor", support_sim="supsim",
                       KL_feature="feat", fc_2="fc", fc_2_wsl="fc", strm="strm", strm_KD="strm",
                       fc_2_sup="klcesup", fc_2_sup_dist="klcesup", fc_2_sup_kl="klcesup",
                       fc_2_sup_dist_cece="klcesup", fc_2_sup_klklcece="klcesup",
                       fc_2_sup_distdistcece="klcesup", fc_2_sup_2="sup2", fc_2_sup_disver="klcesup",
                       fc_2_sup_dist_wsl="klcesup", strm_fc_2_sup_dist="strmsup", strm_1fc_sup="strmsup",
                       fc_1_sup="klsup", fc_sup="klsup", e_dist_1fc_sup="klsup")
    out["labels"] = y
    out["labels20"] = rs.randint(0, w, 20).astype(np.int64)
    for name, ks in recipe_keys.items():
        mk_s, mk_t = keysets[ks]
        s_np, t_np = mk_s(), mk_t()
        lab = out["labels20"] if ks == "supsim" else y
        s_t = {k: t(v, True) for k, v in s_np.items()} if isinstance(s_np, dict) else t(s_np, True)
        t_t = {k: t(v) for k, v in t_np.items()} if isinstance(t_np, dict) else t(t_np)
        d = distillers.Distiller(name, dict(cfg), "cpu")
        res = getattr(d, name)(s_t, t_t, t(lab))
        res["loss"].reshape(()).backward()
        out[f"{name}__loss"] = npy(res["loss"]).reshape(())
        for k, v in res.items():
            if k != "loss" and torch.is_tensor(v):
                out[f"{name}__part__{k}"] = npy(v).reshape(-1)
        if isinstance(s_np, dict):
            for k in s_np:
                out[f"{name}__s__{k}"] = s_np[k]
                g = s_t[k].grad
                out[f"{name}__g__{k}"] = npy(g) if g is not None else np.zeros_like(s_np[k])
        else:
            out[f"{name}__s"] = s_np
            out[f"{name}__g"] = npy(s_t.grad)
        if isinstance(t_np, dict):
            for k in t_np:
                out[f"{name}__t__{k}"] = t_np[k]
        else:
            out[f"{name}__t"] = t_np
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
    print("losses.npz", {k: float(v) for k, v in out.items() if k.endswith("__loss")})


def gen_edist(C):
    """Frame-mean Euclidean heads (model/classifiers/e_dist.py:22-61, COS.py:29-62): outputs + gradients."""
    out = {}
    rs = np.random.RandomState(SEED + 4)
    args = types.SimpleNamespace(seq_len=8, way=5, shot=2)
    sup, s_lab, qry, q_lab = structured_episode(rs, 5, 2, 1, 8, 2048)
    up = rs.standard_normal((5, 5)).astype(np.float32)
    for name, cls in (("edist", C.e_dist), ("cos", C.CosDistance)):
        S, Q = t(sup, True), t(qry, True)
        o = cls(args)(S, t(s_lab), Q)
        lg = o["logits"] if isinstance(o, dict) else o
        (lg * t(up)).sum().backward()
        out.update({f"{name}_logits": npy(lg), f"{name}_grad_support": npy(S.grad), f"{name}_grad_query": npy(Q.grad)})
    out.update(support=sup, support_labels=s_lab, query=qry, upstream=up)
    np.savez_compressed(os.path.join(HERE, "edist.npz"), **out)
    print("edist.npz", out["edist_logits"][0])


def feature_head_inputs(seq_len=8, n_context=2, n_target=1):
    """Seeded inputs of the feature-head fixture; regenerated (not stored) by the tests -- the two Linear
    weights alone are 8 MB.  RandomState streams are stable across numpy versions."""
    rs = np.random.RandomState(SEED + 5)
    fm_c = rs.standard_normal((n_context * seq_len, 512, 7, 7)).astype(np.float32)
    fm_t = rs.standard_normal((n_target * seq_len, 512, 7, 7)).astype(np.float32)
    W = (rs.standard_normal((2, 2048, 512)) * 0.04).astype(np.float32)
    b = (rs.standard_normal((2, 2048)) * 0.1).astype(np.float32)
    up_c = rs.standard_normal((2, n_context, seq_len, 2048)).astype(np.float32)
    up_t = rs.standard_normal((2, n_target, seq_len, 2048)).astype(np.float32)
    return fm_c, fm_t, W, b, up_c, up_t


def gen_feature_heads(root):
    """Student feature heads feeding the path: the reference's own
    model/backbone/resnet18_2fc.py and resnet18_student.py forward + backward, with the ResNet trunk replaced by
    the identity (shim 3: torchvision.models.resnet18(pretrained=True) would download weights), so the input is
    the trunk's [frames, 512, 7, 7] map."""
    import torchvision.models as tvm

    class _Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = torch.nn.Identity(), torch.nn.Identity(), torch.nn.Identity()
    tvm.resnet18 = lambda pretrained=True: _Stub()
    bdir = os.path.join(root, "model", "backbone")

    def load(name):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(bdir, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return getattr(m, name)

    fm_c, fm_t, W, b, up_c, up_t = feature_head_inputs()
    args = types.SimpleNamespace(seq_len=8, num_gpus=1)
    out = {}
    # --- two heads -----------------------------------------------------------------------
    net = load("resnet18_2fc")(args)
    with torch.no_grad():
        net.fc1.weight.copy_(t(W[0])); net.fc1.bias.copy_(t(b[0]))
        net.fc2.weight.copy_(t(W[1])); net.fc2.bias.copy_(t(b[1]))
    C_, T_ = t(fm_c, True), t(fm_t, True)
    cd, td = net(C_, None, T_)
    loss = sum((cd[f"context_features_{h + 1}"] * t(up_c[h])).sum() + (td[f"target_features_{h + 1}"] * t(up_t[h])).sum()
               for h in range(2))
    loss.backward()
    for h in range(2):
        out[f"context_features_{h + 1}"] = npy(cd[f"context_features_{h + 1}"])
        out[f"target_features_{h + 1}"] = npy(td[f"target_features_{h + 1}"])
    gW = np.stack([npy(net.fc1.weight.grad), npy(net.fc2.weight.grad)])
    out.update(grad_bias=np.stack([npy(net.fc1.bias.grad), npy(net.fc2.bias.grad)]),
               grad_weight_rows=gW[:, :32].copy(), grad_weight_norm=np.sqrt((gW.astype(np.float64) ** 2).sum((1, 2))),
               grad_fmap_context_head=npy(C_.grad)[:2].copy(), grad_fmap_target_head=npy(T_.grad)[:2].copy(),
               grad_fmap_context_sum=npy(C_.grad).sum((2, 3)), grad_fmap_target_sum=npy(T_.grad).sum((2, 3)))
    # --- single head (resnet18_student) reuses head 0's parameters ---------------------------
    net1 = load("resnet18_student")(args)
    with torch.no_grad():
        net1.res18_2048.weight.copy_(t(W[0])); net1.res18_2048.bias.copy_(t(b[0]))
    c1, t1 = net1(t(fm_c), None, t(fm_t))
    out.update(student_context=npy(c1), student_target=npy(t1))
    np.savez_compressed(os.path.join(HERE, "feature_heads.npz"), **out)
    print("feature_heads.npz", out["context_features_1"][0, 0, :3], os.path.getsize(os.path.join(HERE, "feature_heads.npz")))


def gen_strm(root):
    """STRM DistanceLoss (model/classifiers/strm_res18_sup.py:162-243) and the strmclassifiers_resnet18_sup wrapper
    (:288-325): the reference's own modules in eval mode (dropout off), forward + every gradient."""
    import model.classifiers.strm_res18_sup as S
    out = {}
    rs = np.random.RandomState(SEED + 9)
    way, shot, L, D, dout = 5, 3, 8, 64, 32
    args = types.SimpleNamespace(seq_len=L, trans_dropout=0.1, trans_linear_out_dim=dout, trans_linear_in_dim=D,
                                 way=way, shot=shot, device="cpu")
    sup, s_lab, qry, q_lab = structured_episode(rs, way, shot, 2, L, D)
    head = S.DistanceLoss(args, 2).eval()
    with torch.no_grad():
        head.clsW.weight.copy_(t(rs.standard_normal(tuple(head.clsW.weight.shape)).astype(np.float32) * 0.15))
        head.clsW.bias.copy_(t(rs.standard_normal(tuple(head.clsW.bias.shape)).astype(np.float32) * 0.1))
    Sg, Qg = t(sup, True), t(qry, True)
    lg = head(Sg, t(s_lab), Qg, "cpu")["logits"]
    up = rs.standard_normal(tuple(lg.shape)).astype(np.float32)
    (lg * t(up)).sum().backward()
    out.update(support=sup, support_labels=s_lab, query=qry, query_labels=q_lab, W=npy(head.clsW.weight),
               b=npy(head.clsW.bias), logits=npy(lg), upstream=up, grad_support=npy(Sg.grad), grad_query=npy(Qg.grad),
               gW=npy(head.clsW.weight.grad), gb=npy(head.clsW.bias.grad))
    # ragged: one class with fewer supports, one class absent
    keep = np.array([i for i, l in enumerate(s_lab) if not (l == 3 or (l == 1 and i % 2 == 0))])
    lg2 = head(t(sup[keep]), t(s_lab[keep]), t(qry), "cpu")["logits"]
    out.update(ragged_keep=keep.astype(np.int64), ragged_logits=npy(lg2))
    np.savez_compressed(os.path.join(HERE, "strm.npz"), **out)


def gen_fusion(T):
    """Teacher MFM fusion forward: the reference's own ThreeTransforTemproal / TwoTransforFusion
    modules and ThreeTRXShiftLoopTime.extract_feature (teacher/code/model.py:1648-1664) in eval(), with the
    deterministic parameters of tests/fusion_fixture.py.  Stores inputs and outputs (the encoders themselves are
    ~0.5 G parameters and are regenerated by the tests)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import fusion_fixture as FF
    args = FF.fusion_args()
    three = T.ThreeTransforTemproal(args).eval()
    two = T.TwoTransforFusion(args).eval()
    FF.fill_parameters(three, "three_fusion")
    FF.fill_parameters(two, "fusion")
    rgb, depth, flow = FF.modality_inputs()
    holder = types.SimpleNamespace(args=args, three_fusion=three, fusion=two)
    with torch.no_grad():
        total = T.ThreeTRXShiftLoopTime.extract_feature(holder, {"rgb": t(rgb), "depth": t(depth), "flow": t(flow)})
        f_three = three.extract_feature(t(rgb), t(depth), t(flow))
        f_two = two.extract_feature(t(rgb), t(depth))
    out = dict(total=npy(total), three=npy(f_three), two_rgb_depth=npy(f_two),
               keys_three=np.array(sorted(three.state_dict().keys())), keys_two=np.array(sorted(two.state_dict().keys())),
               checksum_three=np.float64(sum(float(v.double().sum()) for v in three.state_dict().values())),
               checksum_two=np.float64(sum(float(v.double().sum()) for v in two.state_dict().values())))
    np.savez_compressed(os.path.join(HERE, "fusion.npz"), **out)
    print("fusion.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    distillers, C, T = load_reference(root)
    torch.manual_seed(SEED)
    if len(sys.argv) > 2 and sys.argv[2] == "strm":       # regenerate only the fixtures added in round 2
        gen_strm(root)
        return
    if len(sys.argv) > 2 and sys.argv[2] == "fusion":
        gen_fusion(T)
        return
    gen_fusion(T)
    gen_strm(root)
    gen_otam(T, distillers)
    gen_trx(T, C)
    gen_student(C)
    gen_losses(distillers)
    gen_edist(C)
    gen_feature_heads(root)


if __name__ == "__main__":
    main()
