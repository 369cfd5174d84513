"""Accuracy helper of the reference's utils.py on the lmkd CUDA path (utils.py:116-121)."""
from lmkd import ops


def aggregate_accuracy(test_logits_sample, test_labels):
    """mean(argmax(logits, -1) == labels) as a 0-d float tensor on the logits' device."""
    rows = test_logits_sample.numel() // test_logits_sample.shape[-1]
    return ops.accuracy_count(test_logits_sample, test_labels).float().reshape(()) / rows


def split_first_dim_linear(x, first_two_dims):
    """utils.py helper used by TrxBranch (teacher/code/model.py:1125)."""
    return x.reshape(list(first_two_dims) + list(x.shape[1:]))
