"""CPU-side checks: the C-ABI library loads and exports every symbol include/lmkd.h declares,
the ctypes table covers the header, host-side tables are right, and the product refuses to run
without CUDA instead of falling back."""
import ctypes
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "lmkd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lmkd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lmkd import _ffi
    assert os.path.exists(_ffi.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_ffi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(handle, s)]
    assert not missing, missing
    assert sorted(_ffi.SIGNATURES) == syms, "ctypes table and header disagree"
    assert _ffi.lib().lmkd_version() == 100


def test_workspace_size_queries_run_without_a_gpu():
    from lmkd import _ffi
    lib = _ffi.lib()
    assert lib.lmkd_otam_workspace_bytes(64, 25, 25, 8, 2048, 5) > 64 * 50 * 8 * 2048 * 2
    sh = _ffi.TrxShape(64, 25, 25, 8, 2048, 1152, 3, 5, 5, 0.0, 0, None, 1e-5)
    fwd = lib.lmkd_trx_workspace_bytes(ctypes.byref(sh), 0)
    both = lib.lmkd_trx_workspace_bytes(ctypes.byref(sh), 1)
    assert 0 < fwd < both < 20 * 2 ** 30
    bad = _ffi.TrxShape(1, 5, 5, 8, 2047, 64, 2, 5, 1, 0.0, 0, None, 1e-5)      # D not a multiple of 8
    assert lib.lmkd_trx_workspace_bytes(ctypes.byref(bad), 0) == 0
    assert b"multiples of 8" in lib.lmkd_last_error()
    assert lib.lmkd_sim_pitch(40) == 40 and lib.lmkd_sim_pitch(25) == 32


def test_tuple_tables_match_itertools_and_invert():
    from itertools import combinations
    from lmkd import ops
    for L, c in [(8, 2), (8, 3), (5, 1), (6, 4)]:
        tuples, off, idx = ops.tuple_tables(L, c)
        assert tuples.tolist() == [list(t) for t in combinations(range(L), c)]
        assert off.shape[0] == c * L + 1 and idx.shape[0] == c * tuples.shape[0]
        for j in range(c):
            for l in range(L):
                got = idx[off[j * L + l]:off[j * L + l + 1]].tolist()
                assert got == [t for t, tp in enumerate(tuples.tolist()) if tp[j] == l]


def test_state_dict_keys_match_reference_contract():
    import model.classifiers as C
    a = types.SimpleNamespace(seq_len=8, trans_dropout=0.1, trans_linear_out_dim=16, trans_linear_in_dim=32,
                              way=5, shot=5, temp_set=[2, 3])
    want = ["transformers.pe.pe", "transformers.k_linear.weight", "transformers.k_linear.bias",
            "transformers.v_linear.weight", "transformers.v_linear.bias", "transformers.norm_k.weight",
            "transformers.norm_k.bias", "transformers.norm_v.weight", "transformers.norm_v.bias"]
    for cls in (C.TRX, C.TRX_fixed, C.TRX_2fc, C.TRX_2fcsup, C.TRX_2fcsup_fixed, C.TRX_sup, C.TRX_sup_fixed):
        assert list(cls(a).state_dict().keys()) == want, cls.__name__
    b = C.TrxBranch(a)
    assert "transformers.1.k_linear.weight" in b.state_dict()
    assert b.transformers[1].k_linear.weight.shape == (16, 3 * 32)
    assert C.TRX(a).transformers.pe.pe.shape == (1, 12, 32)


def test_pe_buffer_equals_reference_formula():
    import oracle
    import model.classifiers as C
    pe = C.PositionalEncoding(64, 0.1, max_len=12).pe[0]
    assert torch.equal(pe, oracle.positional_encoding_table(12, 64))


def test_unbuilt_heads_fail_loudly():
    import model.classifiers as C
    with pytest.raises(NotImplementedError):
        C.TRX_1fc_sup                 # named in the reference's __all__, but it has no source there either
    with pytest.raises(AttributeError):
        C.no_such_head
    # every classifier the reference's name2classifier can resolve exists here
    for name in ("TRX", "TRX_fixed", "TRX_2fc", "TRX_2fcsup", "TRX_2fcsup_fixed", "TRX_sup", "TRX_sup_fixed", "CosDistance",
                 "e_dist", "e_dist_fc2", "e_dist_fc2_sup", "e_dist_fc2_sup_fixed", "e_dist_1fc_sup", "strmclassifiers",
                 "strmclassifiers_resnet18", "strmclassifiers_resnet18_sup"):
        assert isinstance(getattr(C, name), type), name


def test_no_cpu_fallback():
    """A CPU tensor must raise, not silently run elsewhere."""
    import distillers
    import model.classifiers as C
    a = types.SimpleNamespace(seq_len=8, trans_dropout=0.0, trans_linear_out_dim=16, trans_linear_in_dim=32,
                              way=5, shot=1, temp_set=[2])
    with pytest.raises(RuntimeError, match="CUDA"):
        C.TRX(a)(torch.zeros(5, 8, 32), torch.arange(5.), torch.zeros(5, 8, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        C.OTAM(a)(torch.zeros(5, 8, 32), torch.arange(5.), torch.zeros(5, 8, 32))
    cfg = dict(temperature=4, soft_loss_weight=2, hard_loss_weight=1, feature_loss_weight=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        distillers.Distiller("KD", cfg, "cpu").KD(torch.zeros(25, 5), torch.zeros(25, 5), torch.zeros(25).long())


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "lite-mkd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)


def test_distiller_has_every_reference_recipe():
    import distillers
    from oracle.losses import RECIPE_NAMES
    for n in RECIPE_NAMES:
        assert callable(getattr(distillers.Distiller, n)), n


def test_episode_generator_shapes_and_determinism():
    from lmkd.episodes import make_episodes
    a = make_episodes(3, 5, 5, 5, 8, 64, teacher_dim=32, modalities=3)
    b = make_episodes(3, 5, 5, 5, 8, 64, teacher_dim=32, modalities=3)
    assert a.support.shape == (3, 25, 8, 64) and a.teacher_query.shape == (3, 25, 8, 32)
    assert a.support_labels.dtype == torch.float32 and a.query_labels.dtype == torch.int64
    assert all(torch.equal(x, y) for x, y in zip(a.tensors(), b.tensors()))
    assert sorted(a.support_labels[0].tolist()) == sorted([float(c) for c in range(5)] * 5)


def test_feature_head_backbones_keep_reference_state_dict_keys_and_have_no_cpu_path():
    """model/backbone/resnet18_2fc.py:31-35 / resnet18_student.py:31-34: `resnet.*`, `fc1.*`, `fc2.*` /
    `res18_2048.*` are the checkpoint keys; the heads run in liblmkd only (CPU tensors raise)."""
    import types

    import torch

    from model.backbone.resnet18_2fc import resnet18_2fc
    from model.backbone.resnet18_student import resnet18_student
    args = types.SimpleNamespace(seq_len=8, num_gpus=1)
    trunk = torch.nn.Sequential(torch.nn.Conv2d(3, 512, 1))
    keys2 = set(resnet18_2fc(args, trunk=trunk).state_dict())
    assert keys2 == {"resnet.0.weight", "resnet.0.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"}
    keys1 = set(resnet18_student(args, trunk=trunk).state_dict())
    assert keys1 == {"resnet.0.weight", "resnet.0.bias", "res18_2048.weight", "res18_2048.bias"}
    assert args.trans_linear_in_dim == 2048
    net = resnet18_2fc(args, trunk=torch.nn.Identity())
    assert net.fc1.weight.shape == (2048, 512)
    with pytest.raises(RuntimeError):
        net(torch.zeros(8, 512, 7, 7), None, torch.zeros(8, 512, 7, 7))
