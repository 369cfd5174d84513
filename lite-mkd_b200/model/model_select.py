"""Plugin boundary of the reference (model/model_select.py) on the lmkd path.

`Student`, `Teacher`, `select_model_student`, `select_model_teacher`, `load_teacher`,
`select_test` keep the reference's names, constructor signature `(args)` and return values, so
trainwandb.py / test.py import them unchanged.  Classifiers come from this package's
`model.classifiers`; backbones are stock PyTorch modules taken from the reference checkout
(see model/backbone/__init__.py) and imported lazily, only for the name that is selected.
"""
import importlib

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

import model.classifiers as classifiers

# args.model_backbone -> (module under model.backbone, class name)   (reference :165-178)
NAME2BACKBONE = {
    "resnet18_student": ("resnet18_student", "resnet18_student"),
    "resnet50_student": ("resnet50_student", "resnet50_stduent"),
    "strm18_student": ("strm18_student", "strm18_student"),
    "resnet18_2fc": ("resnet18_2fc", "resnet18_2fc"),
    "resnet50_2fc": ("resnet50_2fc", "resnet50_2fc"),
    "strmbackbone": ("strmbackbone", "strmbackbone"),
    "meta_baseline": ("meta_baseline", "meta_baseline"),
    "meta_baseline_fc2": ("meta_baseline_fc2", "meta_baseline_fc2"),
    "moblienetv3_fc2": ("moblienetv3", "mobile_large_2fc"),
    "moblienetv3": ("moblienetv3", "mobile_large"),
    "precomputed": ("precomputed", "precomputed"),
}

# args.model_classifier -> class in model.classifiers   (reference :182-199, plus OTAM / TrxBranch)
NAME2CLASSIFIER = {
    "cos": "CosDistance", "TRX": "TRX", "TRX_sup": "TRX_sup", "CTX": "CTX", "TRX_2fc": "TRX_2fc",
    "TRX_1fc_sup": "TRX_1fc_sup", "TRX_2fcsup": "TRX_2fcsup", "TRX_2fcsup_2": "TRX_2fcsup_2",
    "strmclassifiers": "strmclassifiers", "e_dist": "e_dist", "e_dist_fc2": "e_dist_fc2",
    "e_dist_fc2_sup": "e_dist_fc2_sup", "strm_res18": "strmclassifiers_resnet18",
    "strm_res18_sup": "strmclassifiers_resnet18_sup", "strm_1fc_sup": "strm_1fc_sup",
    "e_dist_1fc_sup": "e_dist_1fc_sup", "OTAM": "OTAM", "TrxBranch": "TrxBranch",
}

# args.model_teacher -> class   (reference :222-235)
NAME2TEACHER = {
    "cos": "CosDistance", "e_dist": "e_dist", "e_dist_fc2_sup": "e_dist_fc2_sup_fixed",
    "train_teacher": "TRX", "test_teacher": "TRX_fixed",
    "train_teacher_TRX_sup": "TRX_sup", "test_teacher_TRX_sup_fixed": "TRX_sup_fixed",
    "train_teacher_TRX_2fcsup": "TRX_2fcsup", "test_teacher_TRX_2fcsup_fixed": "TRX_2fcsup_fixed",
    "test_teacher_OTAM": "OTAM", "test_teacher_TrxBranch": "TrxBranch",
}


class Student(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.backbone, self.classifier = select_model_student(args)

    def forward(self, context_feature, context_labels, target_feature):
        context_features, target_features = self.backbone(context_feature, context_labels, target_feature)
        logits = self.classifier(context_features, context_labels, target_features)["logits"]
        return {"logits": logits, "context_features": context_features, "target_features": target_features}


class Teacher(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.train()
        self.args = args
        self.classifier = select_model_teacher(args)

    def forward(self, context_feature, context_labels, target_feature):
        return self.classifier(context_feature, context_labels, target_feature)

    def distribute_model(self):
        # heads run one process per GPU with episodes sharded (lmkd.dist); nothing to wrap here
        return None


def load_teacher(teacher, args):
    """Copy the 9 tensors of the fused teacher's first transformer (key prefix
    'bracnch.transformers.0.', reference :105-117) into `teacher.transformers`."""
    state = torch.load(args.teacher_checkpoint, map_location="cpu")["model_state_dict"]
    pre = "bracnch.transformers.0."
    tr = teacher.transformers
    tr.pe.pe = state[pre + "pe.pe"].to(tr.pe.pe.device)
    for mod, name in ((tr.k_linear, "k_linear"), (tr.v_linear, "v_linear"), (tr.norm_k, "norm_k"),
                      (tr.norm_v, "norm_v")):
        mod.weight = Parameter(state[pre + name + ".weight"].to(mod.weight.device))
        mod.bias = Parameter(state[pre + name + ".bias"].to(mod.bias.device))
    return teacher


def load_student(args):
    student = Student(args)
    state = torch.load(args.test_model_path, map_location="cpu")["model_state_dict"]
    # strip the DataParallel 'module.' level the reference's checkpoints carry (reference :143-151)
    state = {k.replace(".module.", ".", 1) if k.split(".")[2:3] == ["module"] else k: v for k, v in state.items()}
    student.load_state_dict(state)
    return student


def _backbone(name, args):
    mod_name, cls_name = NAME2BACKBONE[name]
    try:
        mod = importlib.import_module("model.backbone." + mod_name)
    except ImportError as e:
        raise ImportError(
            f"backbone '{name}' is stock PyTorch code of the reference and is not part of this build; set "
            "LITE_MKD_REFERENCE=/path/to/Lite-MKD so model.backbone can import it") from e
    return getattr(mod, cls_name)(args)


def select_model_student(args):
    backbone = _backbone(args.model_backbone, args)
    classifier = getattr(classifiers, NAME2CLASSIFIER[args.model_classifier])(args)
    if getattr(args, "num_gpus", 1) > 1 and hasattr(backbone, "resnet"):
        backbone.resnet = torch.nn.DataParallel(backbone.resnet, device_ids=list(range(args.num_gpus)))
    return backbone, classifier


def select_model_teacher(args):
    classifier = getattr(classifiers, NAME2TEACHER[args.model_teacher])(args)
    if args.model_teacher in ("test_teacher", "test_teacher_TRX_sup_fixed"):
        classifier = load_teacher(classifier, args)
    return classifier


def select_test(args):
    teacher = classifiers.TRX_fixed(args)
    return {"teacher": lambda: load_teacher(teacher, args), "student": lambda: load_student(args)}[args.test_model]()
