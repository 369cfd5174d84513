"""GPU side of the teacher-feature store through the C-ABI: the episode gather is a byte
copy (bit-exact against the reference-style per-video loads), the store-fed fused feature-MSE equals the plain
fused pass on the gathered tensor (gradients bit-exact, loss within fp32 summation-order noise) and the
reference's F.mse_loss."""
import random

import numpy as np
import pytest
import torch

import oracle
from lmkd.feature_store import FeatureStore, pack_feature_tree, sample_episode_rows

pytestmark = pytest.mark.gpu


def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def packed(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("teacher_feature"))
    rs = np.random.RandomState(11)
    for c in range(6):
        for v in range(12):
            oracle.write_feature(root, f"c{c}", f"vid{v:02d}", rs.standard_normal((1, 8, 2048)).astype(np.float32))
    out = {}
    for dt in ("fp32", "bf16"):
        path = str(tmp_path_factory.mktemp("store") / f"s_{dt}.lmkd")
        pack_feature_tree(root, path, dt)
        out[dt] = FeatureStore(path)
    return root, out


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_gather_is_the_reference_episode(packed, dtype):
    from lmkd import ops
    d = dev()
    root, stores = packed
    st = stores[dtype]
    dstore = st.to_device(d)
    assert dstore.shape == (72, 8 * 2048)
    _, per_class = oracle.scan_teacher_tree(root)
    for seed in range(3):
        s_ref, _, q_ref, _, _ = oracle.episode_teacher_features(per_class, 5, 5, 5, random.Random(seed))
        s_rows, _, q_rows, _, _ = sample_episode_rows(st, 5, 5, 5, random.Random(seed))
        got_s = ops.episode_gather(dstore, s_rows, 8)
        got_q = ops.episode_gather(dstore, q_rows, 8)
        assert got_s.shape == (25, 8, 2048) and got_s.dtype == torch.float32
        if dtype == "fp32":
            assert torch.equal(got_s.cpu(), s_ref) and torch.equal(got_q.cpu(), q_ref)          # bit-exact
        else:
            assert torch.equal(got_s.cpu(), s_ref.bfloat16().float()) and torch.equal(got_q.cpu(), q_ref.bfloat16().float())
    # batched index tensor [B, N]
    idx = torch.randint(0, 72, (4, 50))
    got = ops.episode_gather(dstore, idx, 8)
    ref = torch.from_numpy(np.array(stores["fp32"].rows))[idx.reshape(-1)].reshape(4, 50, 8, 2048)
    if dtype == "bf16":
        ref = ref.bfloat16().float()
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_store_fed_feature_mse_equals_plain_fused_pass(packed, dtype):
    from lmkd import _ffi, ops
    d = dev()
    _, stores = packed
    dstore = stores[dtype].to_device(d)
    g = torch.Generator().manual_seed(3)
    B, N = 6, 50
    idx = torch.randint(0, 72, (B, N), generator=g)
    student = torch.randn(B, N, 8, 2048, generator=g).to(d)
    teacher = ops.episode_gather(dstore, idx, 8)
    n_e = N * 8 * 2048
    s1 = student.clone().requires_grad_(True)
    l1 = ops.feature_mse(s1, teacher, 1.5, n_e)
    l1.backward()
    s2 = student.clone().requires_grad_(True)
    l2 = ops.feature_mse_from_store(s2, dstore, idx, 1.5, n_e)
    l2.backward()
    assert torch.equal(s1.grad, s2.grad)                                     # same arithmetic per element
    assert abs(l1.item() - l2.item()) <= 1e-5 * abs(l1.item())
    # and the reference's formulation: weight * sum over episodes of F.mse_loss
    ref = 1.5 * sum(torch.nn.functional.mse_loss(student[b].double().cpu(), teacher[b].double().cpu()) for b in range(B))
    assert abs(l2.item() - float(ref)) <= 1e-4 * abs(float(ref))
    _ffi.check_device_status(d)


def test_out_of_range_row_is_reported(packed):
    from lmkd import _ffi, ops
    d = dev()
    _, stores = packed
    dstore = stores["fp32"].to_device(d)
    out = ops.episode_gather(dstore, torch.tensor([0, 72, 3]), 8)
    assert torch.count_nonzero(out[1]) == 0
    with pytest.raises(RuntimeError, match="feature-store index"):
        _ffi.check_device_status(d)
    _ffi.check_device_status(d)          # the flag is cleared by the report


def test_cfg3_sized_gather_and_loss_properties():
    """Config 3 width: 256 episodes x 50 videos from a 4096-video bf16 store (0.86 GB of gathered fp32): gather of a
    permutation is invertible, the store-fed loss of the gathered tensor itself is exactly 0."""
    from lmkd import ops
    d = dev()
    nvid, row = 4096, 8 * 2048
    store = torch.randn(nvid, row, device=d).bfloat16()
    idx = torch.randint(0, nvid, (256, 50), device=d)
    t = ops.episode_gather(store, idx, 8)
    assert torch.equal(t.reshape(-1, row), store[idx.reshape(-1)].float())
    s = t.clone().requires_grad_(True)
    loss = ops.feature_mse_from_store(s, store, idx, 1.0, 50 * row)
    loss.backward()
    assert loss.item() == 0.0 and torch.count_nonzero(s.grad) == 0
    perm = torch.randperm(nvid, device=d)
    back = ops.episode_gather(ops.episode_gather(store, perm, 8).reshape(nvid, row).bfloat16(), torch.argsort(perm), 8)
    assert torch.equal(back.reshape(nvid, row), store.float())
