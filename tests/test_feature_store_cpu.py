"""Host side of the packed teacher-feature store against the oracle restatement of the
reference's per-video loader: same ordering, same bytes, same episode draw.  CPU only (format + index logic;
the device kernels are covered by tests/test_gpu_feature_store.py)."""
import os
import random

import numpy as np
import pytest
import torch

import oracle
from lmkd.feature_store import FeatureStore, pack_feature_tree, sample_episode_rows


def make_tree(root, n_classes=7, L=8, D=2048, seed=3483):
    rs = np.random.RandomState(seed)
    names = [f"class_{c:02d}" for c in rs.permutation(n_classes)]          # creation order != sorted order
    for cname in names:
        for v in rs.permutation(int(rs.randint(9, 14))):
            oracle.write_feature(root, cname, f"v_{cname}_{v:03d}", rs.standard_normal((1, L, D)).astype(np.float32))
    return names


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("teacher_feature"))
    make_tree(root)
    return root


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_pack_preserves_reference_order_and_bytes(tree, tmp_path, dtype):
    out = str(tmp_path / f"store_{dtype}.lmkd")
    hdr = pack_feature_tree(tree, out, dtype)
    st = FeatureStore(out)
    classes, per_class = oracle.scan_teacher_tree(tree)
    assert st.classes == classes and hdr["videos"] == len(st) == sum(len(p) for p in per_class)
    assert (st.L, st.D) == (8, 2048)
    for c, paths in enumerate(per_class):
        rows = st.videos_of_class(c)
        assert len(rows) == len(paths)
        for r, path in zip(rows, paths):
            ref = oracle.load_teacher_feature(path)                 # [1, L, D] fp32, as the reference loads it
            assert st.names[r] == os.path.basename(os.path.dirname(path))
            got = torch.from_numpy(np.array(st.rows[r]))
            if dtype == "fp32":
                assert torch.equal(got.reshape(1, 8, 2048), ref)    # bit-exact
            else:
                assert torch.equal(got.view(torch.bfloat16).reshape(1, 8, 2048), ref.bfloat16())
    assert st.row_of(classes[2], st.names[int(st.videos_of_class(2)[1])]) == int(st.videos_of_class(2)[1])


def test_episode_draw_matches_the_reference_loader(tree, tmp_path):
    """Same random.Random stream -> same classes, same videos, same shuffles, same feature bytes as
    VideoDataset.__getitem__'s teacher-feature path (video_reader.py:403-471)."""
    out = str(tmp_path / "store.lmkd")
    pack_feature_tree(tree, out, "fp32")
    st = FeatureStore(out)
    _, per_class = oracle.scan_teacher_tree(tree)
    for seed in range(5):
        s_ref, sl_ref, q_ref, ql_ref, bc_ref = oracle.episode_teacher_features(per_class, 5, 2, 3, random.Random(seed))
        s_rows, s_lab, q_rows, q_lab, bc = sample_episode_rows(st, 5, 2, 3, random.Random(seed))
        assert bc == bc_ref
        assert torch.equal(s_lab, sl_ref) and torch.equal(q_lab, ql_ref)
        got_s = torch.from_numpy(np.array(st.rows[s_rows.numpy()])).reshape(-1, 8, 2048)
        got_q = torch.from_numpy(np.array(st.rows[q_rows.numpy()])).reshape(-1, 8, 2048)
        assert torch.equal(got_s, s_ref) and torch.equal(got_q, q_ref)
        assert s_ref.shape == (10, 8, 2048) and q_ref.shape == (15, 8, 2048)


def test_bad_inputs_fail_loudly(tree, tmp_path):
    with pytest.raises(ValueError):
        pack_feature_tree(tree, str(tmp_path / "x"), "fp16")
    bad = tmp_path / "not_a_store"
    bad.write_bytes(b"garbage" * 10)
    with pytest.raises(RuntimeError):
        FeatureStore(str(bad))
    empty = tmp_path / "empty_tree"
    empty.mkdir()
    with pytest.raises(RuntimeError):
        pack_feature_tree(str(empty), str(tmp_path / "y"))
    from lmkd import ops
    with pytest.raises(RuntimeError):                                   # no CPU path for the gather
        ops.episode_gather(torch.zeros(4, 16), torch.zeros(2, dtype=torch.int64), 2)
