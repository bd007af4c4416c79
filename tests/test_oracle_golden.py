"""Pin the CPU oracle (oracle/dformer_oracle.py) to golden vectors produced by the
unmodified reference (oracle/make_golden.py).  CPU only."""
import json
import os

import pytest
import torch

from golden_util import make_inputs, make_state
from oracle import dformer_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _layout(variant):
    with open(os.path.join(G, "state_dict_layouts.json")) as f:
        return json.load(f)[variant]


@pytest.mark.parametrize("aten", [False, True])
def test_tiny_eval_forward_matches_reference(aten):
    g = torch.load(os.path.join(G, "tiny_eval_64x96.pt"))
    lay = _layout("DFormer-Tiny")
    P = make_state(lay["shapes"], seed=g["seed"])
    B, H, W = g["size"]
    rgb, hha, label, bases = make_inputs(B, H, W, 40, seed=g["seed"])
    v = O.VARIANTS["DFormer-Tiny"]
    with torch.no_grad(), O.aten_faithful(aten):
        r = O.forward(P, rgb, hha, bases, v["dims"], v["depths"], label=label, training=False, return_all=True)
    for a, b in zip(r["outs"], g["outs"]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r["small"], g["small"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r["out"], g["out"].float(), rtol=2e-3, atol=2e-3)   # golden stored as fp16
    torch.testing.assert_close(r["loss"], g["loss"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("aten", [False, True])
def test_tiny_train_forward_backward_matches_reference(aten):
    """aten=True: the ATen-faithful mode (the ops the reference dispatches; the autocast yardstick) is pinned by the same vectors."""
    g = torch.load(os.path.join(G, "tiny_train_96x128.pt"))
    lay = _layout("DFormer-Tiny")
    P = make_state(lay["shapes"], seed=g["seed"])
    for k in lay["trainable"]:
        P[k].requires_grad_(True)
    B, H, W = g["size"]
    rgb, hha, label, bases = make_inputs(B, H, W, 40, seed=g["seed"])
    v = O.VARIANTS["DFormer-Tiny"]
    stats = {}
    with O.aten_faithful(aten):
        r = O.forward(P, rgb, hha, bases, v["dims"], v["depths"], label=label, training=True,
                      new_stats=stats, return_all=True)
    for a, b in zip(r["outs"], g["outs"]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r["small"], g["small"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r["out"].mean(dim=(2, 3)), g["out_mean"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r["loss"], g["loss"], rtol=1e-5, atol=1e-5)
    r["loss"].backward()
    missing = sorted(k for k in lay["trainable"] if P[k].grad is None)
    assert missing == g["no_grad"]                       # the unused stem_e_fc1/2
    for k, n in g["grad_norm"].items():
        assert abs(P[k].grad.norm().item() - n) <= 1e-3 * n + 1e-7, k
    for k, gr in g["grads"].items():
        cos = torch.nn.functional.cosine_similarity(P[k].grad.flatten(), gr.flatten(), dim=0)
        assert cos > 0.99999, (k, cos)
    for k, s in g["new_stats"].items():
        torch.testing.assert_close(stats[k], s, rtol=1e-4, atol=1e-5)


def test_nmf2d_matches_reference():
    g = torch.load(os.path.join(G, "nmf2d_512x64_8x10.pt"))
    x = g["x"].reshape(2, 512, 80)
    torch.testing.assert_close(O.nmf2d(x, g["bases"], 7).reshape(2, 512, 8, 10), g["eval"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(O.nmf2d(x, g["bases"], 6).reshape(2, 512, 8, 10), g["train"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("size_in,size_out", [((7, 7), (60, 80)), ((15, 20), (60, 80)), ((8, 12), (64, 96)), ((3, 4), (12, 16))])
def test_bilinear_restatement_matches_torch(size_in, size_out):
    x = torch.randn(2, 5, *size_in)
    ref = torch.nn.functional.interpolate(x, size_out, mode="bilinear", align_corners=False)
    torch.testing.assert_close(O.bilinear_resize_nchw(x, size_out), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("hw", [(60, 80), (30, 40), (15, 20), (2, 3), (7, 7)])
def test_adaptive_pool_restatement_matches_torch(hw):
    x = torch.randn(2, 6, *hw)
    torch.testing.assert_close(O.adaptive_avg_pool_7(x), torch.nn.functional.adaptive_avg_pool2d(x, (7, 7)), rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference_small_variant():
    """Live cross-check on DFormer-Small (a variant that has no committed golden)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(G), "..", "oracle"))
    from oracle import make_golden as MG
    m = MG.build_reference("DFormer-Small", 40)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert {k: list(v) for k, v in shapes.items()} == _layout("DFormer-Small")["shapes"]
    P = make_state(shapes, seed=3)
    m.load_state_dict(P, strict=True)
    m.eval()
    rgb, hha, label, bases = make_inputs(1, 64, 64, 40, seed=3)
    with torch.no_grad():
        outs, small, out, loss = MG.ref_forward(m, rgb, hha, label, bases)
        v = O.VARIANTS["DFormer-Small"]
        r = O.forward(P, rgb, hha, bases, v["dims"], v["depths"], label=label, return_all=True)
    torch.testing.assert_close(r["out"], out, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(r["loss"], loss, rtol=1e-5, atol=1e-5)


def test_trainpre_oracle_matches_reference_golden():
    """oracle/trainpre_oracle.py (numpy restatement of TrainPre incl. OpenCV's fixed-point resize) is pinned bit-exactly to the
    outputs of the unmodified reference preprocessing run with real cv2 (oracle/make_golden_trainpre.py)."""
    import random

    import numpy as np

    from oracle import trainpre_oracle as T
    g = torch.load(os.path.join(G, "trainpre.pt"))
    assert len(g["cases"]) >= 12
    for c in g["cases"]:
        random.seed(c["seed"])
        r, l, m = T.train_pre(c["rgb"].numpy(), c["gt"].numpy(), c["modal"].numpy(), g["mean"], g["std"], c["scales"], c["crop"], c["sign"])
        assert torch.from_numpy(np.ascontiguousarray(r)).float().equal(c["out_rgb"]), c["seed"]
        assert torch.from_numpy(np.ascontiguousarray(l)).equal(c["out_gt"]), c["seed"]
        assert torch.from_numpy(np.ascontiguousarray(m)).float().equal(c["out_modal"]), c["seed"]
