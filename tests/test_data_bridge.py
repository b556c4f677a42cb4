"""the obj36 h5 reader (xggm_b200.data.Obj36Reader) on in-memory stand-ins of the reference's files
(h5py is absent in the build image; the real-file path is exercised when it is importable)."""
import json
import os

import numpy as np
import pytest
import torch


def _fake(n_img, seed=0):
    rng = np.random.default_rng(seed)
    obj, info, adj = {}, {}, {}
    for k in range(n_img):
        img_id = f"COCO_train2014_{k:012d}"
        h, w = int(rng.integers(200, 640)), int(rng.integers(200, 640))
        xy = np.sort(rng.random((36, 2, 2)).astype(np.float32), axis=-1)
        boxes = np.stack([xy[:, 0, 0] * w, xy[:, 1, 0] * h, xy[:, 0, 1] * w, xy[:, 1, 1] * h], axis=-1).astype(np.float32)
        obj[img_id] = {"features": np.maximum(rng.standard_normal((36, 2048)), 0).astype(np.float32), "boxes": boxes}
        info[img_id] = {"img_id": img_id, "img_h": h, "img_w": w, "num_boxes": 36}
        c = rng.random((36, 36)).astype(np.float32)
        adj[img_id] = (c + c.T) / (c + c.T).max()
    return obj, info, adj


def test_batch_matches_the_reference_getitem_arithmetic():
    from xggm_b200.data import Obj36Reader
    obj, info, adj = _fake(5)
    ids = list(obj)[::-1]
    feats, boxes, a = Obj36Reader(obj, info, adj, pin=False).batch(ids)
    assert feats.shape == (5, 36, 2048) and boxes.shape == (5, 36, 4) and a.shape == (5, 36, 36)
    assert feats.dtype == boxes.dtype == a.dtype == torch.float32
    for i, img_id in enumerate(ids):
        b = obj[img_id]["boxes"].copy()                 # src/vqa/vqacpv2_data.py:112-117, line by line
        b[:, (0, 2)] /= info[img_id]["img_w"]
        b[:, (1, 3)] /= info[img_id]["img_h"]
        assert np.array_equal(boxes[i].numpy(), b)
        assert np.array_equal(feats[i].numpy(), obj[img_id]["features"])
        assert np.array_equal(a[i].numpy(), adj[img_id])
    assert float(boxes.min()) >= 0 and float(boxes.max()) <= 1 + 1e-5


def test_rejects_bad_groups_and_splits_without_adjacency():
    from xggm_b200.data import Obj36Reader
    obj, info, adj = _fake(2)
    ids = list(obj)
    f, b, a = Obj36Reader(obj, info, None, pin=False).batch(ids)
    assert a is None
    obj[ids[0]]["boxes"] = obj[ids[0]]["boxes"] * 10.0          # outside the image
    with pytest.raises(ValueError, match="outside"):
        Obj36Reader(obj, info, adj, pin=False).batch(ids)
    info[ids[1]]["num_boxes"] = 35
    with pytest.raises(ValueError, match="expected 36"):
        Obj36Reader(obj, info, adj, pin=False).batch([ids[1]])


def test_open_reads_real_h5_files(tmp_path):
    h5py = pytest.importorskip("h5py")
    from xggm_b200.data import Obj36Reader
    obj, info, adj = _fake(3)
    with h5py.File(tmp_path / "train_obj36.h5", "w") as f:
        for k, v in obj.items():
            g = f.create_group(k)
            g.create_dataset("features", data=v["features"])
            g.create_dataset("boxes", data=v["boxes"])
    with h5py.File(tmp_path / "train_obj36_adj_v2.h5", "w") as f:
        for k, v in adj.items():
            f.create_dataset(name=k, data=v, dtype=np.float32)
    json.dump(list(info.values()), open(tmp_path / "train_obj36_info.json", "w"))
    feats, boxes, a = Obj36Reader.open(str(tmp_path), "train").batch(list(obj))
    assert np.array_equal(feats[1].numpy(), obj[list(obj)[1]]["features"]) and a.shape == (3, 36, 36)


def test_open_without_h5py_fails_loudly():
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py present")
    except ImportError:
        pass
    from xggm_b200.data import Obj36Reader
    with pytest.raises(RuntimeError, match="needs h5py"):
        Obj36Reader.open("/nonexistent", "train")
