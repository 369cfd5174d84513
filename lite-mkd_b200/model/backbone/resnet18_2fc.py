"""`resnet18_2fc` backbone with the reference's interface (model/backbone/resnet18_2fc.py:16-80): a ResNet-18
trunk followed by two Linear(512 -> 2048) heads; returns the two-head feature dicts the *_2fc classifiers take.
State-dict keys match the reference (`resnet.*`, `fc1.*`, `fc2.*`)."""
import torch.nn as nn

from ._feature_heads import PooledLinearHeads, make_trunk


class resnet18_2fc(nn.Module):
    def __init__(self, args, trunk=None):
        super().__init__()
        self.args = args
        self.args.trans_linear_in_dim = 2048
        self.num_patches = 16
        self.adap_max = nn.AdaptiveMaxPool2d((4, 4))        # parameter-free; kept for attribute parity
        self.resnet = trunk if trunk is not None else make_trunk("resnet18")
        heads = PooledLinearHeads(("fc1", "fc2"), 512, 2048, out_hw=4)
        self.fc1, self.fc2 = heads.layers["fc1"], heads.layers["fc2"]
        object.__setattr__(self, "_heads", heads)            # not a registered child: keys stay fc1.* / fc2.*

    def forward(self, context_feature, context_labels, target_feature):
        ctx, tgt = self._heads(self.resnet(context_feature), self.resnet(target_feature), self.args.seq_len)
        return ({"context_features_1": ctx[0], "context_features_2": ctx[1]},
                {"target_features_1": tgt[0], "target_features_2": tgt[1]})

    def distribute_model(self):
        if getattr(self.args, "num_gpus", 1) > 1:
            self.resnet = nn.DataParallel(self.resnet, device_ids=list(range(self.args.num_gpus)))
