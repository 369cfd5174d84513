"""Oracle restatement of the teacher-feature side of the reference's episode loader (test infrastructure, see
oracle/__init__.py).  SURVEY.md §8f rank 2.  Citations are reference paths (file:line)."""
from __future__ import annotations

import os
import random

import numpy as np
import torch


def write_feature(root: str, class_name: str, video_name: str, feature: np.ndarray) -> str:
    """teacher/code/extract_multi_feature.py:116-121 -- the extractor saves `feature.cpu().numpy()`
    ([1, L, 2048] fp32) as <root>/<class>/<video>/feature.npy."""
    save_path = os.path.join(root, class_name, video_name)
    os.makedirs(save_path, exist_ok=True)
    np.save(os.path.join(save_path, "feature.npy"), feature)
    return save_path


def scan_teacher_tree(root: str):
    """video_reader.py:252-268 -- sorted class folders, sorted video folders, first file of each video folder;
    class id = index of the class folder.  Returns per-class lists of feature paths."""
    class_folders = sorted(os.listdir(root))
    per_class = []
    for class_folder in class_folders:
        paths = []
        for video_folder in sorted(os.listdir(os.path.join(root, class_folder))):
            feature = os.listdir(os.path.join(root, class_folder, video_folder))[0]
            paths.append(os.path.join(root, class_folder, video_folder, feature))
        per_class.append(paths)
    return class_folders, per_class


def load_teacher_feature(path: str) -> torch.Tensor:
    """video_reader.py:393-394 -- np.load + torch.from_numpy of one video's [1, L, 2048] array."""
    return torch.from_numpy(np.load(path))


def episode_teacher_features(per_class, way, shot, n_queries, rng: random.Random):
    """video_reader.py:403-461, teacher-feature side only: sample `way` classes, shot + n_queries videos of each,
    load every video's feature, shuffle supports and queries, concatenate (torch.cat, :470-471)."""
    classes = list(range(len(per_class)))
    batch_classes = rng.sample(classes, way)
    sup, qry = [], []
    for bl, bc in enumerate(batch_classes):
        n_total = len(per_class[bc])
        idxs = rng.sample([i for i in range(n_total)], shot + n_queries)
        for idx in idxs[0:shot]:
            sup.append((load_teacher_feature(per_class[bc][idx]), bl))
        for idx in idxs[shot:]:
            qry.append((load_teacher_feature(per_class[bc][idx]), bl))
    rng.shuffle(sup)
    rng.shuffle(qry)
    s_feat, s_lab = zip(*sup)
    q_feat, q_lab = zip(*qry)
    return (torch.cat(s_feat), torch.FloatTensor(s_lab), torch.cat(q_feat), torch.FloatTensor(q_lab), batch_classes)
