"""GPU input pipeline (dformer_b200.data.TrainPre, csrc/data.cu) -- bit-exact against (a) the golden outputs of the unmodified
reference TrainPre run with real cv2 and (b) the numpy oracle on full-size NYUDepthv2-shaped batches."""
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_train_pre_matches_reference_golden_bit_exactly():
    from dformer_b200.data import TrainPre
    g = torch.load(os.path.join(G, "trainpre.pt"))
    for c in g["cases"]:
        cfg = SimpleNamespace(train_scale_array=c["scales"], image_height=c["crop"][0], image_width=c["crop"][1])
        pre = TrainPre(g["mean"], g["std"], sign=c["sign"], config=cfg)
        random.seed(c["seed"])                                   # same `random` stream as the reference consumed
        rgb, gt, modal = pre(c["rgb"][None].cuda(), c["gt"][None].cuda(), c["modal"][None].cuda())
        assert rgb[0].cpu().equal(c["out_rgb"]), c["seed"]
        assert modal[0].cpu().equal(c["out_modal"]), c["seed"]
        assert gt[0].cpu().equal(c["out_gt"].long()), c["seed"]


def test_train_pre_full_size_batch_matches_oracle():
    from oracle import trainpre_oracle as T

    from dformer_b200.data import TrainPre
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    scales = [0.5, 0.75, 1, 1.25, 1.5, 1.75]
    B, H, W = 6, 480, 640
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    modal = rng.integers(0, 256, (B, H, W, 1), dtype=np.uint8).repeat(3, axis=3)
    gt = rng.integers(0, 41, (B, H, W), dtype=np.uint8)
    cfg = SimpleNamespace(train_scale_array=scales, image_height=480, image_width=640)
    pre = TrainPre(mean, std, config=cfg)
    random.seed(11)
    o_rgb, o_gt, o_modal = pre(torch.from_numpy(rgb).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(modal).cuda())
    random.seed(11)                                              # the batch draws per sample in order, like B dataset items would
    for b in range(B):
        r, l, m = T.train_pre(rgb[b], gt[b], modal[b], mean, std, scales, (480, 640))
        assert o_rgb[b].cpu().equal(torch.from_numpy(np.ascontiguousarray(r)).float()), b
        assert o_modal[b].cpu().equal(torch.from_numpy(np.ascontiguousarray(m)).float()), b
        assert o_gt[b].cpu().equal(torch.from_numpy(np.ascontiguousarray(l)).long()), b


def test_val_pre_matches_reference_formula():
    """ValPre (dataloader.py:112-122): float64 normalisation (depth always with 0.48 / 0.28), HWC -> CHW; labels untouched."""
    from dformer_b200.data import ValPre
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (2, 37, 53, 3), dtype=np.uint8)
    modal = rng.integers(0, 256, (2, 37, 53, 3), dtype=np.uint8)
    gt = rng.integers(0, 41, (2, 37, 53), dtype=np.uint8)
    o_rgb, o_gt, o_modal = ValPre(mean, std)(torch.from_numpy(rgb).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(modal).cuda())
    norm = lambda im, m, s: ((im.astype(np.float64) / 255.0 - np.asarray(m)) / np.asarray(s)).transpose(0, 3, 1, 2)
    assert o_rgb.cpu().equal(torch.from_numpy(np.ascontiguousarray(norm(rgb, mean, std))).float())
    assert o_modal.cpu().equal(torch.from_numpy(np.ascontiguousarray(norm(modal, [0.48] * 3, [0.28] * 3))).float())
    assert o_gt.cpu().equal(torch.from_numpy(gt).long())
