"""Bridge from the reference's on-disk obj36 format to the block's input tensors.

The reference trainers read, per image id, a group of ``{split}_obj36.h5`` with ``features [36,2048] f32`` and
``boxes [36,4] f32`` (written by data/preprocess/vqa/tsv2h5.py), image sizes from ``{split}_obj36_info.json``, and
the object-similarity adjacency ``[36,36] f32`` from ``{split}_obj36_adj_v2.h5`` (a dataset per image id, written
by data/preprocess/vqa/compute_adjacency.py:94-96); ``VQATorchDataset.__getitem__`` normalises the boxes to [0,1]
(src/vqa/vqacpv2_data.py:104-125).  ``Obj36Reader`` does the same per BATCH of image ids and returns pinned host
tensors ready for ``GraphedStep.prefetch`` / a non-blocking H2D copy:

    reader = Obj36Reader.open(root, "train")                       # needs h5py (absent in the build image)
    feats, boxes, adj = reader.batch(img_ids)                      # [B,36,2048], [B,36,4], [B,36,36] fp32, pinned

The storage is any mapping with h5py's indexing protocol (``store[str(img_id)]["features"][:]``), so the logic is
testable without h5py (tests/test_data_bridge.py uses dictionaries of numpy arrays).
"""
import json
import os

import numpy as np
import torch

N_OBJ, FEAT_DIM = 36, 2048


class Obj36Reader:
    def __init__(self, obj_store, obj_info, adj_store=None, pin=None):
        """obj_store[img_id] -> {"features": [36,2048], "boxes": [36,4]}; obj_info[img_id] -> {"img_h", "img_w",
        "num_boxes"}; adj_store[img_id] -> [36,36] (only the train / dev_test splits have one,
        src/vqa/vqacpv2_data.py:76-80)."""
        self.obj, self.info, self.adj = obj_store, obj_info, adj_store
        self.pin = torch.cuda.is_available() if pin is None else pin

    @classmethod
    def open(cls, root, split, with_adj=None):
        """Open ``{root}/{split}_obj36.h5`` (+ ``_info.json``, + ``_adj_v2.h5``) as the reference does
        (src/vqa/vqacpv2_data.py:69-80)."""
        try:
            import h5py
        except ImportError as e:  # the build image has no h5py: fail loudly, there is no silent fallback
            raise RuntimeError("xggm_b200.data.Obj36Reader.open needs h5py to read the reference's *_obj36.h5 files") from e
        obj = h5py.File(os.path.join(root, f"{split}_obj36.h5"), "r")
        with open(os.path.join(root, f"{split}_obj36_info.json")) as f:
            info = {d["img_id"]: d for d in json.load(f)}
        if with_adj is None:
            with_adj = split in ("train", "dev_test")
        adj = h5py.File(os.path.join(root, f"{split}_obj36_adj_v2.h5"), "r") if with_adj else None
        return cls(obj, info, adj)

    def _alloc(self, *shape):
        t = torch.empty(*shape, dtype=torch.float32)
        return t.pin_memory() if self.pin else t

    def batch(self, img_ids):
        """(feats [B,36,2048], boxes [B,36,4] normalised to [0,1], adj [B,36,36] or None) for a list of image ids."""
        B = len(img_ids)
        feats, boxes = self._alloc(B, N_OBJ, FEAT_DIM), self._alloc(B, N_OBJ, 4)
        adj = self._alloc(B, N_OBJ, N_OBJ) if self.adj is not None else None
        fn, bn = feats.numpy(), boxes.numpy()
        an = adj.numpy() if adj is not None else None
        for i, img_id in enumerate(img_ids):
            grp, meta = self.obj[f"{img_id}"], self.info[img_id]
            f = np.asarray(grp["features"][:], dtype=np.float32)
            b = np.asarray(grp["boxes"][:], dtype=np.float32).copy()
            if not (meta["num_boxes"] == len(b) == len(f) == N_OBJ) or f.shape[1] != FEAT_DIM:
                raise ValueError(f"image {img_id}: expected {N_OBJ} boxes x {FEAT_DIM} features, got {f.shape} / {b.shape}")
            b[:, (0, 2)] /= meta["img_w"]          # vqacpv2_data.py:113-115
            b[:, (1, 3)] /= meta["img_h"]
            if not ((b < 1 + 1e-5).all() and (-b < 1e-5).all()):   # the reference's assert_array_less pair, :116-117
                raise ValueError(f"image {img_id}: normalised boxes outside [0, 1]")
            fn[i], bn[i] = f, b
            if an is not None:
                a = np.asarray(self.adj[f"{img_id}"][:], dtype=np.float32)
                if a.shape != (N_OBJ, N_OBJ):
                    raise ValueError(f"image {img_id}: adjacency has shape {a.shape}")
                an[i] = a
        return feats, boxes, adj
