"""GPU parity of the student feature heads feeding the path through the C-ABI:
patch pooling + fc1 / fc2 against the reference's own resnet18_2fc / resnet18_student outputs
(tests/golden/feature_heads.npz, trunk = identity) and against the oracle at config-2 size.
Pooling is fp32 (exact maxima, mean within fp32 rounding); the Linear layers are bf16 contractions with fp32
accumulation: outputs rel 1e-2, gradients rel-L2 2e-2 (north_star tolerances)."""
import importlib.util
import os
import types

import numpy as np
import pytest
import torch

import oracle

from conftest import assert_close, rel_l2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def T(x, grad=False, device=None):
    t = torch.from_numpy(np.asarray(x)).clone()
    if device is not None:
        t = t.to(device)
    return t.requires_grad_(grad)


def inputs():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.feature_head_inputs()


def load_heads(net, W, b, d):
    with torch.no_grad():
        for h, name in enumerate(("fc1", "fc2")):
            getattr(net, name).weight.copy_(T(W[h], device=d))
            getattr(net, name).bias.copy_(T(b[h], device=d))


def test_two_head_backbone_vs_reference():
    from model.backbone.resnet18_2fc import resnet18_2fc
    d = dev()
    z = np.load(os.path.join(G, "feature_heads.npz"))
    fm_c, fm_t, W, b, up_c, up_t = inputs()
    net = resnet18_2fc(types.SimpleNamespace(seq_len=8, num_gpus=1), trunk=torch.nn.Identity()).to(d)
    load_heads(net, W, b, d)
    C_, T_ = T(fm_c, True, d), T(fm_t, True, d)
    cd, td = net(C_, None, T_)
    loss = 0
    for h in range(2):
        c, t = cd[f"context_features_{h + 1}"], td[f"target_features_{h + 1}"]
        assert c.shape == (2, 8, 2048) and t.shape == (1, 8, 2048)
        assert rel_l2(c, z[f"context_features_{h + 1}"]) < 1e-2
        assert rel_l2(t, z[f"target_features_{h + 1}"]) < 1e-2
        assert_close(c.detach().cpu().numpy(), z[f"context_features_{h + 1}"], rtol=1e-2, atol=2e-2)
        loss = loss + (c * T(up_c[h], device=d)).sum() + (t * T(up_t[h], device=d)).sum()
    loss.backward()
    gb = torch.stack([net.fc1.bias.grad, net.fc2.bias.grad])
    assert_close(gb.cpu().numpy(), z["grad_bias"], rtol=1e-4, atol=1e-3)     # fp32 column sums
    gW = torch.stack([net.fc1.weight.grad, net.fc2.weight.grad])
    assert rel_l2(gW[:, :32], z["grad_weight_rows"]) < 1e-2
    assert_close(gW.double().pow(2).sum((1, 2)).sqrt().cpu().numpy(), z["grad_weight_norm"], rtol=1e-2)
    assert rel_l2(C_.grad[:2], z["grad_fmap_context_head"]) < 1e-2
    assert rel_l2(T_.grad[:2], z["grad_fmap_target_head"]) < 1e-2
    assert rel_l2(C_.grad.sum((2, 3)), z["grad_fmap_context_sum"]) < 1e-2
    # the pooling backward only ever touches window maxima: same sparsity pattern as the reference's
    nz_ref = z["grad_fmap_context_head"] != 0
    assert np.array_equal(C_.grad[:2].cpu().numpy() != 0, nz_ref)


def test_single_head_backbone_vs_reference():
    from model.backbone.resnet18_student import resnet18_student
    d = dev()
    z = np.load(os.path.join(G, "feature_heads.npz"))
    fm_c, fm_t, W, b, _, _ = inputs()
    net = resnet18_student(types.SimpleNamespace(seq_len=8, num_gpus=1), trunk=torch.nn.Identity()).to(d)
    with torch.no_grad():
        net.res18_2048.weight.copy_(T(W[0], device=d))
        net.res18_2048.bias.copy_(T(b[0], device=d))
    c, t = net(T(fm_c, device=d), None, T(fm_t, device=d))
    assert rel_l2(c, z["student_context"]) < 1e-2 and rel_l2(t, z["student_target"]) < 1e-2


@pytest.mark.parametrize("geom", [(7, 7, 4), (8, 6, 3), (5, 5, 5), (9, 10, 4), (4, 4, 1)])
def test_frame_pool_fwd_bwd_vs_oracle(geom):
    from lmkd import ops
    d = dev()
    H, W, o = geom
    g = torch.Generator().manual_seed(31 * H + W)
    x = torch.randn(37, 24, H, W, generator=g)
    up = torch.randn(37, 24, generator=g)
    xo = x.clone().requires_grad_(True)
    ref = oracle.frame_pool(xo, o)
    (ref * up).sum().backward()
    xg = x.to(d).requires_grad_(True)
    got = ops.frame_pool(xg, o)
    torch.testing.assert_close(got.cpu(), ref.detach(), rtol=1e-6, atol=1e-6)
    (got * up.to(d)).sum().backward()
    torch.testing.assert_close(xg.grad.cpu(), xo.grad, rtol=1e-6, atol=1e-7)


def test_frame_pool_rejects_maps_beyond_its_slab():
    from lmkd import ops
    with pytest.raises(RuntimeError, match="frame_pool"):
        ops.frame_pool(torch.zeros(2, 4, 14, 14, device=dev()), 4)


@pytest.mark.parametrize("shape", [(50, 512, 2048, 2), (200, 512, 2048, 1), (133, 64, 200, 3), (25600, 512, 2048, 2)])
def test_feature_heads_vs_oracle(shape):
    """Ragged row counts, odd head counts, and the config-2 size (64 episodes x 50 videos x 8 frames)."""
    from lmkd import ops
    d = dev()
    rows, din, dout, heads = shape
    g = torch.Generator().manual_seed(rows + heads)
    x = torch.randn(rows, din, generator=g)
    W = torch.randn(heads, dout, din, generator=g) * din ** -0.5
    b = torch.randn(heads, dout, generator=g) * 0.1
    up = torch.randn(heads, rows, dout, generator=g)
    xo, Wo, bo = x.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.stack([xo @ Wo[h].t() + bo[h] for h in range(heads)])
    (ref * up).sum().backward()
    xg, Wg, bg = (v.to(d).requires_grad_(True) for v in (x, W, b))
    got = ops.feature_heads(xg, Wg, bg)
    assert got.shape == (heads, rows, dout)
    assert rel_l2(got, ref) < 1e-2
    (got * up.to(d)).sum().backward()
    assert rel_l2(xg.grad, xo.grad) < 1e-2
    assert rel_l2(Wg.grad, Wo.grad) < 1e-2
    assert rel_l2(bg.grad, bo.grad) < 1e-4          # fp32 sums of the fp32 upstream gradient


def test_feature_heads_feed_the_two_head_classifier():
    """End to end across the widened boundary: pooled 512-d frame features -> fc1 / fc2 -> TRX_2fc logits,
    gradients reach the Linear weights (train_task shape, trainwandb.py:205-232)."""
    import model.classifiers as Cm
    from model.backbone.resnet18_2fc import resnet18_2fc
    d = dev()
    args = types.SimpleNamespace(seq_len=8, num_gpus=1, trans_dropout=0.0, trans_linear_out_dim=128, way=5, shot=1,
                                 temp_set=[2], trans_linear_in_dim=2048)
    net = resnet18_2fc(args, trunk=torch.nn.Identity()).to(d)
    head = Cm.TRX_2fc(args).to(d).eval()
    g = torch.Generator().manual_seed(5)
    fm_c = torch.randn(5 * 8, 512, 7, 7, generator=g).to(d)
    fm_t = torch.randn(10 * 8, 512, 7, 7, generator=g).to(d)
    labels = torch.arange(5, dtype=torch.float32, device=d)
    cd, td = net(fm_c, labels, fm_t)
    out = head(cd, labels, td)["logits"]
    lg = out["fc_1"] if isinstance(out, dict) else out
    assert lg.shape[-2:] == (10, 5) and torch.isfinite(lg).all()
    lg.sum().backward()
    assert net.fc1.weight.grad is not None and torch.isfinite(net.fc1.weight.grad).all()
    assert net.fc1.weight.grad.abs().sum() > 0
