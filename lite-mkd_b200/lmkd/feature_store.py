"""Packed teacher-feature store.

The reference keeps the teacher's multi-modal features as one ``feature.npy`` of shape [1, L, 2048] (fp32) per
video under ``<teacher_path>/<class>/<video>/`` (writer: teacher/code/extract_multi_feature.py:113-121) and the
student's loader does one ``np.load`` per video per episode (video_reader.py:388-395), 50 small reads for a
5-way 5-shot task.  Here the tree is packed ONCE into a single file

    b"LMKDFS01" | u64 header bytes | JSON header | zero pad to 4096 | rows [videos, L*D] fp32 or bf16

which is memory-mapped, uploaded to HBM as one tensor, and from then on an episode is a list of row indices:
``ops.episode_gather`` materialises [.., L, D] fp32, ``ops.feature_mse_from_store`` feeds the fused D2M loss
straight from the store.  Classes and videos are ordered as the reference orders them: sorted listings
(video_reader.py:252-268), class id = position of the class folder.
"""
from __future__ import annotations

import json
import os
import random

import numpy as np
import torch

MAGIC = b"LMKDFS01"
_ALIGN = 4096
_DTYPES = {"fp32": (np.float32, torch.float32), "bf16": (np.uint16, torch.bfloat16)}


def _scan(root: str):
    """(class_names, [(class_id, video_name, path)]) in the reference's order (video_reader.py:252-268)."""
    classes = sorted(os.listdir(root))
    videos = []
    for cid, cname in enumerate(classes):
        cdir = os.path.join(root, cname)
        if not os.path.isdir(cdir):
            raise RuntimeError(f"{cdir} is not a class folder")
        for vname in sorted(os.listdir(cdir)):
            files = os.listdir(os.path.join(cdir, vname))
            if not files:
                raise RuntimeError(f"{os.path.join(cdir, vname)} holds no feature file")
            videos.append((cid, vname, os.path.join(cdir, vname, files[0])))     # the loader takes entry [0], :266
    return classes, videos


def pack_feature_tree(root: str, out_path: str, dtype: str = "fp32") -> dict:
    """Pack ``<root>/<class>/<video>/feature.npy`` into one store file; returns the header."""
    if dtype not in _DTYPES:
        raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
    classes, videos = _scan(root)
    if not videos:
        raise RuntimeError(f"no videos under {root}")
    first = np.load(videos[0][2])
    if first.ndim != 3 or first.shape[0] != 1:
        raise RuntimeError(f"{videos[0][2]}: expected a [1, L, D] array, got {first.shape}")
    L, D = int(first.shape[1]), int(first.shape[2])
    if (L * D) % 8:
        raise RuntimeError("L*D must be a multiple of 8")
    header = {"version": 1, "L": L, "D": D, "dtype": dtype, "videos": len(videos), "classes": classes,
              "class_of": [v[0] for v in videos], "names": [v[1] for v in videos]}
    blob = json.dumps(header).encode()
    data_off = -(-(len(MAGIC) + 8 + len(blob)) // _ALIGN) * _ALIGN
    with open(out_path, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(len(blob)).tobytes())
        f.write(blob)
        f.write(b"\0" * (data_off - f.tell()))
        for _, _, path in videos:
            a = np.load(path)
            if a.shape != (1, L, D):
                raise RuntimeError(f"{path}: shape {a.shape} differs from [1, {L}, {D}]")
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
            if dtype == "bf16":      # round to nearest even, exactly what tensor.bfloat16() does
                a = torch.from_numpy(a).bfloat16().view(torch.int16).numpy().view(np.uint16)
            f.write(a.tobytes())
    return header


class FeatureStore:
    """Memory-mapped view of a packed store; ``to_device`` puts all rows in HBM."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise RuntimeError(f"{path} is not an lmkd feature store")
            n = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            self.header = json.loads(f.read(n).decode())
        h = self.header
        self.L, self.D, self.dtype = h["L"], h["D"], h["dtype"]
        self.classes, self.names = h["classes"], h["names"]
        self.class_of = np.asarray(h["class_of"], dtype=np.int64)
        off = -(-(len(MAGIC) + 8 + n) // _ALIGN) * _ALIGN
        self.rows = np.memmap(path, dtype=_DTYPES[self.dtype][0], mode="r", offset=off,
                              shape=(h["videos"], self.L * self.D))
        self._by_class = [np.nonzero(self.class_of == c)[0] for c in range(len(self.classes))]

    def __len__(self):
        return self.rows.shape[0]

    def videos_of_class(self, c: int) -> np.ndarray:
        """Store rows of class c, in the order the reference's ``get_rand_vid(label, idx)`` indexes them."""
        return self._by_class[c]

    def row_of(self, class_name: str, video_name: str) -> int:
        c = self.classes.index(class_name)
        for r in self._by_class[c]:
            if self.names[r] == video_name:
                return int(r)
        raise KeyError((class_name, video_name))

    def to_device(self, device) -> torch.Tensor:
        """[videos, L*D] tensor (fp32 or bf16) on `device`; one upload, then every episode is index work."""
        tdt = _DTYPES[self.dtype][1]
        out = torch.empty(self.rows.shape, dtype=tdt, device=device)
        step = max(1, (256 << 20) // (self.rows.shape[1] * out.element_size()))      # ~256 MB per copy
        for r0 in range(0, len(self), step):
            block = np.array(self.rows[r0:r0 + step])                               # private, writable copy
            if self.dtype == "bf16":      # stored as uint16 bit patterns; torch has no uint16 on older versions
                host = torch.from_numpy(block.view(np.int16)).view(torch.bfloat16)
            else:
                host = torch.from_numpy(block)
            out[r0:r0 + step].copy_(host)
        return out


def sample_episode_rows(store: FeatureStore, way: int, shot: int, n_queries: int, rng: random.Random,
                        classes=None):
    """Row indices and labels of one task, drawn the way ``VideoDataset.__getitem__`` draws them
    (video_reader.py:403-461): `way` classes by random.sample, shot + n_queries distinct videos per class,
    supports and queries shuffled independently.  Returns (support_rows, support_labels, query_rows, query_labels,
    batch_classes)."""
    pool = list(range(len(store.classes))) if classes is None else list(classes)
    batch_classes = rng.sample(pool, way)
    sup, qry = [], []
    for bl, bc in enumerate(batch_classes):
        vids = store.videos_of_class(bc)
        idxs = rng.sample(list(range(len(vids))), shot + n_queries)
        sup += [(int(vids[i]), bl) for i in idxs[:shot]]
        qry += [(int(vids[i]), bl) for i in idxs[shot:]]
    rng.shuffle(sup)
    rng.shuffle(qry)
    s_rows, s_lab = zip(*sup)
    q_rows, q_lab = zip(*qry)
    return (torch.tensor(s_rows, dtype=torch.int64), torch.tensor(s_lab, dtype=torch.float32),
            torch.tensor(q_rows, dtype=torch.int64), torch.tensor(q_lab, dtype=torch.float32), batch_classes)
