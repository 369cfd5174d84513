"""Teacher multi-modal fusion forward: the oracle restatement against the fixture made from
the reference's own modules (tests/golden/make_golden.py::gen_fusion), and the state_dict key contract of the
product modules.  CPU only; parameters are regenerated from tests/fusion_fixture.py."""
import os

import numpy as np
import pytest
import torch

import fusion_fixture as FF

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def modules():
    import model.fusion as MF
    args = FF.fusion_args()
    three, two = MF.ThreeTransforTemproal(args).eval(), MF.TwoTransforFusion(args).eval()
    FF.fill_parameters(three, "three_fusion")
    FF.fill_parameters(two, "fusion")
    return three, two


def test_state_dict_keys_and_parameter_stream_match_the_reference(modules):
    z = np.load(os.path.join(G, "fusion.npz"))
    three, two = modules
    assert sorted(three.state_dict().keys()) == list(z["keys_three"])
    assert sorted(two.state_dict().keys()) == list(z["keys_two"])
    for m, key in ((three, "checksum_three"), (two, "checksum_two")):
        got = sum(float(v.double().sum()) for v in m.state_dict().values())
        assert abs(got - float(z[key])) <= 1e-6 * max(1.0, abs(float(z[key])))


def test_oracle_fusion_matches_reference_outputs(modules):
    import oracle.fusion as OF
    z = np.load(os.path.join(G, "fusion.npz"))
    three, two = modules
    rgb, depth, flow = (torch.from_numpy(x) for x in FF.modality_inputs())
    with torch.no_grad():
        f3 = OF.fusion_encoder([rgb, depth, flow], three.state_dict(), 3, FF.TRANS_NUM)
        f2 = OF.fusion_encoder([rgb, depth], two.state_dict(), 2, FF.TRANS_NUM)
        tot = OF.mfm_extract_feature(rgb, depth, flow, three.state_dict(), two.state_dict(), FF.TRANS_NUM, FF.SHIFT)
    for got, key in ((f3, "three"), (f2, "two_rgb_depth"), (tot, "total")):
        ref = torch.from_numpy(z[key])
        assert (got - ref).abs().max().item() <= 2e-4 * ref.abs().max().item(), key


def test_fusion_modules_refuse_train_mode_and_cpu_tensors():
    import model.fusion as MF
    args = FF.fusion_args()
    args.trans_num = 0                     # parameter containers only: no encoder layers to allocate
    m = MF.TwoTransforFusion(args)
    x = torch.zeros(1, 8, 2048)
    with pytest.raises(RuntimeError, match="inference only"):
        m.extract_feature(x, x)
