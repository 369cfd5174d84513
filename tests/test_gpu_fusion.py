"""GPU parity of the teacher multi-modal fusion forward through the C ABI against the fixture
made from the reference's own ThreeTransforTemproal / TwoTransforFusion / ThreeTRXShiftLoopTime.extract_feature.

Tolerance: every Linear is a bf16 contraction with fp32 accumulate (K up to 6144), LayerNorm / softmax / residuals
in fp32 -> features rel-L2 <= 1e-2 and max |err| <= 3e-2 of the output scale (the north_star's 1e-2 for bf16 paths)."""
import os

import numpy as np
import pytest
import torch

import fusion_fixture as FF
from conftest import record_error, rel_l2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def fusion():
    import model.fusion as MF
    d = dev()
    m = MF.MultiModalFusion(FF.fusion_args()).eval()
    FF.fill_parameters(m.three_fusion, "three_fusion")
    FF.fill_parameters(m.fusion, "fusion")
    return m.to(d)


def close(got, ref, what):
    ref = torch.from_numpy(ref)
    got = got.detach().float().cpu()
    r = rel_l2(got, ref, what)
    m = (got - ref).abs().max().item() / ref.abs().max().item()
    record_error("test_gpu_fusion", **{what + ".max_over_scale": m})
    assert r < 1e-2 and m < 3e-2, (what, r, m)


def test_fusion_encoders_and_extract_feature_vs_reference(fusion):
    z = np.load(os.path.join(G, "fusion.npz"))
    d = dev()
    rgb, depth, flow = (torch.from_numpy(x).to(d) for x in FF.modality_inputs())
    close(fusion.three_fusion.extract_feature(rgb, depth, flow), z["three"], "fusion_three")
    close(fusion.fusion.extract_feature(rgb, depth), z["two_rgb_depth"], "fusion_two")
    total = fusion.extract_feature({"rgb": rgb.cpu(), "depth": depth.cpu(), "flow": flow.cpu()})
    close(total, z["total"], "fusion_total")
    assert total.shape == (FF.N_VIDEOS, 8, 2048) and total.is_cuda


def test_fusion_batch_rows_are_independent_and_shift_is_a_roll(fusion):
    """Videos do not interact: a 70-video batch (ragged last GEMM tile) reproduces the 6-video rows; a rolled input
    with shift 0 equals the un-rolled input with the kernel's own shift."""
    d = dev()
    rgb, depth, _ = (torch.from_numpy(x).to(d) for x in FF.modality_inputs())
    base = fusion.fusion.extract_feature(rgb, depth)
    big_r, big_d = rgb.repeat(12, 1, 1)[:70].contiguous(), depth.repeat(12, 1, 1)[:70].contiguous()
    big = fusion.fusion.extract_feature(big_r, big_d)
    assert rel_l2(big[:6], base) < 1e-5 and rel_l2(big[66:70], base[:4]) < 1e-5
    rolled = torch.cat((depth[:, 1:], depth[:, :1]), dim=1).contiguous()
    a = fusion.fusion._run([rgb, rolled])
    b = fusion.fusion._run([rgb, depth], shifts=[0, 1])
    assert rel_l2(a, b) < 1e-6


def test_fusion_rejects_bad_shapes(fusion):
    d = dev()
    x = torch.zeros(2, 8, 2048, device=d)
    with pytest.raises(RuntimeError):
        fusion.fusion.extract_feature(x, torch.zeros(2, 8, 1024, device=d))
    with pytest.raises(RuntimeError):
        fusion.fusion.extract_feature(torch.zeros(2, 9, 2048, device=d), torch.zeros(2, 9, 2048, device=d))


def test_fusion_weight_pack_follows_parameter_updates(fusion):
    """The bf16 weight pack is cached per module and rebuilt when a parameter is written in place (optimizer step,
    load_state_dict): the output must change with the weights and come back with them."""
    d = dev()
    rgb, depth, _ = (torch.from_numpy(x).to(d) for x in FF.modality_inputs())
    two = fusion.fusion
    base = two.extract_feature(rgb, depth).clone()
    pack0 = two._pack
    assert two._packed() is pack0                      # unchanged parameters: same pack
    saved = two.f1.weight.detach().clone()
    with torch.no_grad():
        two.f1.weight.mul_(0.5)
    halved = two.extract_feature(rgb, depth)
    assert two._pack is not pack0
    bias = two.f1.bias.detach()
    assert rel_l2(halved - bias, 0.5 * (base - bias)) < 1e-2
    with torch.no_grad():
        two.f1.weight.copy_(saved)
    assert rel_l2(two.extract_feature(rgb, depth), base) < 1e-6
