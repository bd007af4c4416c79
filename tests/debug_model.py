"""Stage-by-stage parity dump of the CUDA model against the CPU/GPU oracle (debug aid, GPU box; lives under tests/ because it executes the oracle).
usage: python tests/debug_model.py [variant] [B] [H] [W] [fp32|bf16] [train|eval]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from types import SimpleNamespace  # noqa: E402

import torch.nn as nn  # noqa: E402
from golden_util import make_inputs, make_state  # noqa: E402
from oracle import dformer_oracle as O  # noqa: E402

from dformer_b200 import EncoderDecoder  # noqa: E402


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "DFormer-Tiny"
    B, H, W = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((2, 2), (3, 96), (4, 128)))
    prec = sys.argv[5] if len(sys.argv) > 5 else "fp32"
    mode = sys.argv[6] if len(sys.argv) > 6 else "train"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ncls = 40
    cfg = SimpleNamespace(backbone=variant, decoder="ham", decoder_embed_dim=512, num_classes=ncls, drop_path_rate=0.0, aux_rate=0.0,
                          device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision=prec)
    m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    P = make_state(shapes, seed=2)
    m.load_state_dict(P, strict=True)
    m.cuda()
    m.train(mode == "train")
    m.decode_head.dropout = None
    rgb, hha, label, bases = make_inputs(B, H, W, ncls, seed=2)
    m.decode_head.injected_bases = bases.cuda()
    dev = "cuda"
    Pd = {k: v.to(dev) for k, v in P.items()}
    trainable = [k for k, p in m.named_parameters() if p.requires_grad and not k.startswith("encoder_backbone.stem_e_fc")]
    for k in trainable:
        Pd[k].requires_grad_(True)
    v = O.VARIANTS[variant]
    r = O.forward(Pd, rgb.to(dev), hha.to(dev), bases.to(dev), v["dims"], v["depths"], label=label.to(dev), training=(mode == "train"), return_all=True)
    feats, small = m._small_logits(rgb.to(dev), hha.to(dev))
    loss, out = m._upsample(small, (H, W), label.to(dev))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(feats, r["outs"])):
        print(f"stage{i} out  rel={rel(a, b)[0]:.3e} max={rel(a, b)[1]:.3e}  |ref|max={b.abs().max().item():.3f}")
    print(f"small logits rel={rel(small, r['small'])[0]:.3e} max={rel(small, r['small'])[1]:.3e}")
    print(f"out          rel={rel(out, r['out'])[0]:.3e} max={rel(out, r['out'])[1]:.3e}")
    print(f"loss {loss.item():.6f} ref {r['loss'].item():.6f}")
    agree = (out.argmax(1) == r["out"].argmax(1)).float().mean().item()
    print(f"argmax agreement {agree * 100:.3f}%")
    if mode == "train":
        loss.backward()
        r["loss"].backward()
        torch.cuda.synchronize()
        rows = []
        named = dict(m.named_parameters())
        for k in trainable:
            g, gr = named[k].grad, Pd[k].grad
            if g is None:
                rows.append((9.0, k, "MISSING", 0, 0))
                continue
            cos = torch.nn.functional.cosine_similarity(g.flatten().float(), gr.flatten(), dim=0).item()
            rows.append((1 - cos, k, f"cos={cos:.6f}", rel(g, gr)[0], gr.norm().item()))
        rows.sort(reverse=True)
        print("worst gradients:")
        for r_ in rows[:25]:
            print(f"  {r_[1]:70s} {r_[2]} rel={r_[3]:.3e} |ref|={r_[4]:.3e}")
        print("min cos over params:", 1 - rows[0][0], " n_params:", len(rows))


if __name__ == "__main__":
    main()
