"""Generate tests/golden/* by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
It imports the reference modules through the mmcv/mmengine stand-ins in oracle/ref_shim,
loads the deterministic weights of tests/golden_util.py into them and records outputs,
loss and gradients.  The composition encoder -> head -> interpolate -> criterion follows
models/builder.py:193-208,230 with the fork's `(outs, None)` tuple unwrapped.
"""
import contextlib
import io
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shim"), "/root/reference", os.path.join(ROOT, "tests")]
from golden_util import make_inputs, make_state  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_reference(backbone, num_classes, syncbn=False):
    with contextlib.redirect_stdout(io.StringIO()):
        import logging
        logging.disable(logging.CRITICAL)
        from models.builder import EncoderDecoder
        cfg = SimpleNamespace(backbone=backbone, decoder="ham", decoder_embed_dim=512, num_classes=num_classes,
                              drop_path_rate=0.0, aux_rate=0.0, device="cpu", pretrained_model=None,
                              bn_eps=1e-3, bn_momentum=0.1, background=255)
        m = EncoderDecoder(cfg, criterion=nn.CrossEntropyLoss(reduction="none", ignore_index=255),
                           norm_layer=nn.BatchNorm2d, syncbn=syncbn)
    return m


def ref_forward(m, rgb, hha, label, bases):
    """builder.py:198-203,230; NMF bases injected by patching the CPU torch.rand call site
    (ham_head.py:111) so the same draw can be replayed elsewhere."""
    ham = m.decode_head.hamburger.ham
    ham._build_bases = lambda B, S, D, R, cuda=False: F.normalize(bases.clone(), dim=1)
    outs, _ = m.encoder_backbone(rgb, hha)
    small = m.decode_head(outs)
    out = F.interpolate(small, size=rgb.shape[-2:], mode="bilinear", align_corners=False)
    loss = None
    if label is not None:
        loss = m.criterion(out, label.long())[label.long() != m.cfg.background].mean()
    return outs, small, out, loss


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    # ---- (C) state_dict layouts of every variant: pins names + shapes
    layouts = {}
    for name, ncls in (("DFormer-Tiny", 40), ("DFormer-Small", 40), ("DFormer-Base", 37), ("DFormer-Large", 40)):
        m = build_reference(name, ncls)
        layouts[name] = {"num_classes": ncls,
                         "shapes": {k: list(v.shape) for k, v in m.state_dict().items()},
                         "trainable": [k for k, p in m.named_parameters() if p.requires_grad],
                         "n_params": sum(p.numel() for p in m.parameters())}
    with open(os.path.join(OUT, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f)

    # ---- (A) Tiny, eval forward, 64x96, B=1
    m = build_reference("DFormer-Tiny", 40)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(make_state(shapes, seed=1), strict=True)
    m.eval()
    rgb, hha, label, bases = make_inputs(1, 64, 96, 40, seed=1)
    with torch.no_grad():
        outs, small, out, loss = ref_forward(m, rgb, hha, label, bases)
    torch.save({"variant": "DFormer-Tiny", "num_classes": 40, "seed": 1, "size": (1, 64, 96),
                "outs": [o.clone() for o in outs], "small": small, "out": out.half(), "loss": loss},
               os.path.join(OUT, "tiny_eval_64x96.pt"))

    # ---- (B) Tiny, train-mode forward + backward, 96x128, B=2
    m = build_reference("DFormer-Tiny", 40)
    m.load_state_dict(make_state(shapes, seed=2), strict=True)
    m.train()
    m.decode_head.dropout = None                       # Dropout2d off for parity (SURVEY 8c RNG iii)
    rgb, hha, label, bases = make_inputs(2, 96, 128, 40, seed=2)
    outs, small, out, loss = ref_forward(m, rgb, hha, label, bases)
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    keep = [k for i, k in enumerate(sorted(grads)) if i % 29 == 0 or "layer_scale" in k and i % 7 == 0]
    torch.save({"variant": "DFormer-Tiny", "num_classes": 40, "seed": 2, "size": (2, 96, 128),
                "outs": [o.detach().clone() for o in outs], "small": small.detach(), "loss": loss.detach(),
                "out_mean": out.detach().mean(dim=(2, 3)),
                "grad_norm": {k: g.norm().item() for k, g in grads.items()},
                "grads": {k: grads[k].clone() for k in keep},
                "no_grad": sorted(k for k, p in m.named_parameters() if p.grad is None),
                "new_stats": {k: v.clone() for k, v in m.state_dict().items()
                              if k.endswith(("running_mean", "running_var", "num_batches_tracked"))}},
               os.path.join(OUT, "tiny_train_96x128.pt"))

    # ---- (D) NMF2D stand-alone, eval (7 steps) and train (6 steps), D=512 R=64 N=8x10
    from models.decoders.ham_head import NMF2D
    with contextlib.redirect_stdout(io.StringIO()):
        nmf = NMF2D({"device": "cpu"})
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 512, 8, 10, generator=g)
    b = torch.rand(2, 512, 64, generator=g)
    nmf._build_bases = lambda B, S, D, R, cuda=False: F.normalize(b.clone(), dim=1)
    res = {"x": x, "bases": b}
    for mode in ("eval", "train"):
        nmf.train(mode == "train")
        with torch.no_grad():
            res[mode] = nmf(x)
    torch.save(res, os.path.join(OUT, "nmf2d_512x64_8x10.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
