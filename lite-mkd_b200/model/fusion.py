"""Teacher multi-modal fusion forward on the lmkd path.

Reference: teacher/code/model.py -- TrainablePositionalEncoding (:1135-1151), TwoTransforFusion (:1300-1331),
ThreeTransforTemproal (:1361-1392) and ThreeTRXShiftLoopTime.extract_feature (:1648-1664), which
teacher/code/extract_multi_feature.py:113-121 runs in eval() to write the `feature.npy` files the student's D2M
feature loss is trained against.  Same class names, constructor signature `(args)`, attribute names and state_dict
keys as the reference (the parameter containers are torch's own nn.TransformerEncoder / nn.Linear / nn.Embedding /
nn.LayerNorm), so a teacher checkpoint loads unchanged; the arithmetic runs in liblmkd.so (ops.fusion_encoder_forward).

Inference only, like the extraction program: calling a module in train() mode raises (the reference's training
forward applies three dropouts; the teacher training program is out of scope)."""
import torch
import torch.nn as nn

from lmkd import ops

IN_CHANNELS = 2048          # hard-coded by the reference (model.py:1304, :1364)


class TrainablePositionalEncoding(nn.Module):
    """LayerNorm(x + position_embeddings[0..L)) then dropout (teacher/code/model.py:1135-1151); evaluated inside the
    fusion kernels, this class only owns the parameters."""

    def __init__(self, max_position_embeddings, hidden_size, dropout=0.1):
        super().__init__()
        self.position_embeddings = nn.Embedding(max_position_embeddings, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size)
        self.dropout = nn.Dropout(dropout)


class _FusionEncoder(nn.Module):
    N_MOD = 0

    def __init__(self, args, out_channels=None, dropout=0.1):
        super().__init__()
        n = self.N_MOD
        for i in range(n):
            setattr(self, f"positionEncoding{i + 1}", TrainablePositionalEncoding(args.seq_len, IN_CHANNELS))
        encoder_layer = nn.TransformerEncoderLayer(d_model=IN_CHANNELS * n, nhead=n, batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(encoder_layer, num_layers=args.trans_num)
        self.f1 = nn.Linear(IN_CHANNELS * n, IN_CHANNELS)
        self.dropout = nn.Dropout(dropout)
        self._pack = None
        self._pack_key = None

    def _packed(self):
        """bf16 weight pack, rebuilt when any parameter was written (in-place update, load_state_dict, .to())."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._pack is None or key != self._pack_key:
            pes = [getattr(self, f"positionEncoding{i + 1}") for i in range(self.N_MOD)]
            self._pack = ops.PackedFusionEncoder(pes, list(self.transformer_encoder.layers), self.f1, self.N_MOD)
            self._pack_key = key
        return self._pack

    def _run(self, xs, shifts=None, out=None, accumulate=False):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: inference only on the lmkd path -- call .eval() first "
                               "(feature extraction runs in eval mode, teacher/code/extract_multi_feature.py:114)")
        with torch.no_grad():
            return ops.fusion_encoder_forward(self._packed(), xs, shifts, out=out, accumulate=accumulate)


class ThreeTransforTemproal(_FusionEncoder):
    """Three-modality encoder, d_model 6144, 3 heads (teacher/code/model.py:1361-1392)."""
    N_MOD = 3

    def forward(self, x1, y1, z1):
        return self._run([x1, y1, z1])

    def extract_feature(self, x1, y1, z1):
        return self._run([x1, y1, z1])


class TwoTransforFusion(_FusionEncoder):
    """Two-modality encoder, d_model 4096, 2 heads (teacher/code/model.py:1300-1331); positionEncoding1 is shared by
    both first-modality inputs of `forward`, positionEncoding2 by both second-modality inputs."""
    N_MOD = 2

    def forward(self, x1, x2, y1, y2):
        return self._run([x1, y1]), self._run([x2, y2])

    def extract_feature(self, x1, y1):
        return self._run([x1, y1])


class MultiModalFusion(nn.Module):
    """The fusion part of ThreeTRXShiftLoopTime (teacher/code/model.py:1586-1664): attributes `three_fusion` and
    `fusion` as in the reference; `extract_feature(feature)` takes the dict {'rgb', 'depth', 'flow'} of per-modality
    frame features and returns the summed [videos, seq_len, 2048] multi-modal feature."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.fusion = TwoTransforFusion(args)
        self.three_fusion = ThreeTransforTemproal(args)

    def extract_feature(self, feature):
        a = self.args
        dev = self.f_device()
        rgb, depth, flow = (feature[k].reshape(-1, a.seq_len, a.trans_linear_in_dim).to(dev)
                            for k in ("rgb", "depth", "flow"))
        s = int(a.shirt_num)
        out = self.three_fusion._run([rgb, depth, flow])
        # the rolled copies of :1654-1662 are not materialised: the positional-encoding kernel reads frame (l + s) % L
        self.fusion._run([rgb, depth], shifts=[0, s], out=out, accumulate=True)
        self.fusion._run([rgb, flow], shifts=[0, s], out=out, accumulate=True)
        return out

    def f_device(self):
        return self.fusion.f1.weight.device
