"""Host-side mirror of the reference API (no GPU): state_dict layout, init equivalence, loud failure on CPU."""
import json
import os
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

G = os.path.join(os.path.dirname(__file__), "golden")


def _cfg(name, ncls):
    return SimpleNamespace(backbone=name, decoder="ham", decoder_embed_dim=512, num_classes=ncls, drop_path_rate=0.1, aux_rate=0.0,
                           device="cpu", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255)


@pytest.mark.parametrize("name", ["DFormer-Tiny", "DFormer-Small", "DFormer-Base", "DFormer-Large"])
def test_state_dict_layout_identical_to_reference(name):
    from dformer_b200 import EncoderDecoder
    lay = json.load(open(os.path.join(G, "state_dict_layouts.json")))[name]
    m = EncoderDecoder(_cfg(name, lay["num_classes"]), norm_layer=nn.BatchNorm2d)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert list(got) == list(lay["shapes"])                      # same keys, same order
    assert got == lay["shapes"]
    assert sum(p.numel() for p in m.parameters()) == lay["n_params"]
    assert [k for k, p in m.named_parameters() if p.requires_grad] == lay["trainable"]
    # every parameter except the fork's unused stem_e_fc1/2 has a slot in the gradient arena, exactly once
    enc = m.encoder_backbone._build_plan().layout
    assert sorted(set(k for k, _ in m.encoder_backbone.named_parameters()) - set(enc.order)) == [
        "stem_e_fc1.bias", "stem_e_fc1.weight", "stem_e_fc2.bias", "stem_e_fc2.weight"]
    assert len(enc.order) == len(set(enc.order))
    head = m.decode_head._build_plan().layout
    assert set(head.order) == set(k for k, _ in m.decode_head.named_parameters())
    offs = sorted((s.offset, s.offset + s.numel) for s in enc.slots.values())
    assert all(a[1] <= b[0] for a, b in zip(offs, offs[1:]))     # no overlap


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree only exists in the build container")
def test_same_seed_gives_the_reference_initialisation():
    from oracle import make_golden as MG
    from dformer_b200 import EncoderDecoder
    torch.manual_seed(123)
    ref = MG.build_reference("DFormer-Tiny", 40)
    torch.manual_seed(123)
    cfg = _cfg("DFormer-Tiny", 40)
    cfg.drop_path_rate = 0.0
    mine = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert mine.decode_head.squeeze.bn.eps == 1e-3 and ref.decode_head.squeeze.bn.eps == 1e-3


def test_no_cpu_fallback():
    from dformer_b200 import EncoderDecoder
    m = EncoderDecoder(_cfg("DFormer-Tiny", 40), norm_layer=nn.BatchNorm2d).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.randn(1, 3, 64, 64), torch.randn(1, 3, 64, 64))
