"""`resnet18_student` backbone with the reference's interface (model/backbone/resnet18_student.py:16-76): ResNet-18
trunk + one Linear(512 -> 2048) (`res18_2048`); returns ([Ns, L, 2048], [Nq, L, 2048])."""
import torch.nn as nn

from ._feature_heads import PooledLinearHeads, make_trunk


class resnet18_student(nn.Module):
    def __init__(self, args, trunk=None):
        super().__init__()
        self.args = args
        self.args.trans_linear_in_dim = 2048
        self.num_patches = 16
        self.adap_max = nn.AdaptiveMaxPool2d((4, 4))
        self.resnet = trunk if trunk is not None else make_trunk("resnet18")
        heads = PooledLinearHeads(("res18_2048",), 512, 2048, out_hw=4)
        self.res18_2048 = heads.layers["res18_2048"]
        object.__setattr__(self, "_heads", heads)

    def forward(self, context_feature, context_labels, target_feature):
        ctx, tgt = self._heads(self.resnet(context_feature), self.resnet(target_feature), self.args.seq_len)
        return ctx[0], tgt[0]

    def distribute_model(self):
        if getattr(self.args, "num_gpus", 1) > 1:
            self.resnet = nn.DataParallel(self.resnet, device_ids=list(range(self.args.num_gpus)))
