"""Row N1 host logic: the fused optimizer's state dict is `torch.optim.AdamW`'s (reference numbering: `group_weight`,
utils/init_func.py:26-70) and checkpoints keep the reference's dict format (utils/engine/engine.py:101-186)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn


def _model(variant="DFormer-Tiny"):
    from dformer_b200 import EncoderDecoder
    cfg = SimpleNamespace(backbone=variant, decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.0, aux_rate=0.0,
                          device="cpu", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="fp32")
    return EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d)


def _group_weight(module):
    """The reference rule restated for the test (utils/init_func.py:26-70)."""
    decay, no_decay = [], []
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            decay.append(m.weight)
            if m.bias is not None:
                no_decay.append(m.bias)
        elif isinstance(m, (nn.BatchNorm2d, nn.LayerNorm, nn.GroupNorm)):
            no_decay += [m.weight, m.bias]
    return [dict(params=decay, lr=1.0), dict(params=no_decay, weight_decay=0.0, lr=1.0)]


def test_fused_adamw_state_dict_is_the_reference_optimizers():
    from dformer_b200.optim import FusedAdamW
    torch.manual_seed(0)
    m = _model()
    opt = FusedAdamW(m, lr=3e-4, weight_decay=0.02)
    assert opt.state_dict()["state"] == {}                                   # nothing stepped yet, like torch
    for st in opt.state:
        st["m"].normal_()
        st["v"].uniform_()
    opt.step_count = 7
    sd = opt.state_dict()
    groups = _group_weight(m)
    stock = torch.optim.AdamW(groups, lr=1.0, weight_decay=0.5)
    assert [g["params"] for g in sd["param_groups"]] == [g["params"] for g in stock.state_dict()["param_groups"]]
    assert set(sd["param_groups"][0]) == set(stock.state_dict()["param_groups"][0])
    stock.load_state_dict(sd)
    assert stock.param_groups[0]["lr"] == 3e-4 and stock.param_groups[0]["weight_decay"] == 0.02 and stock.param_groups[1]["weight_decay"] == 0.0
    params = groups[0]["params"] + groups[1]["params"]
    names = {id(p): k for k, p in m.named_parameters()}
    with_state = {names[id(params[i])] for i in sd["state"]}
    # every Linear / conv / BatchNorm parameter on the hot path has moments; the unused stem_e_fc layers never get a gradient
    assert not any("stem_e_fc" in k or "layer_scale" in k or ".norm" in k for k in with_state)
    assert len(with_state) == len(params) - 4
    for i, e in sd["state"].items():
        assert e["exp_avg"].shape == params[i].shape and float(e["step"]) == 7.0
    # and back: a fresh fused optimizer restored from the stock optimizer's state holds the same moments
    opt2 = FusedAdamW(_model(), lr=1.0, weight_decay=0.0)
    opt2.load_state_dict(stock.state_dict())
    assert (opt2.step_count, opt2.lr, opt2.weight_decay) == (7, 3e-4, 0.02)
    for a, b in zip(opt.state, opt2.state):
        live = a["lrm"] != 0
        assert torch.equal(a["m"] * live, b["m"]) and torch.equal(a["v"] * live, b["v"]) and torch.equal(a["wd"], b["wd"])


def test_checkpoint_file_has_the_reference_format(tmp_path):
    from dformer_b200.engine import restore_checkpoint, save_checkpoint
    from dformer_b200.optim import FusedAdamW
    torch.manual_seed(1)
    m = _model()
    opt = FusedAdamW(m)
    path = str(tmp_path / "epoch-9.pt")
    save_checkpoint(path, nn.ModuleDict({"module": m}), opt, epoch=9, iteration=4321)      # DDP-style "module." prefix is stripped
    ck = torch.load(path, weights_only=False)
    assert list(ck) == ["model", "optimizer", "epoch", "iteration"] and (ck["epoch"], ck["iteration"]) == (9, 4321)
    assert list(ck["model"]) == list(m.state_dict())
    torch.manual_seed(2)
    m2 = _model()
    assert restore_checkpoint(path, m2, FusedAdamW(m2)) == (10, 4321)                       # epoch to continue with (engine.py:178)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    # the reference's restore view (keys with "module.") is accepted; a foreign optimizer state is refused
    torch.save({"model": {"module." + k: v for k, v in ck["model"].items()}, "optimizer": ck["optimizer"], "epoch": 0, "iteration": 0}, path)
    assert restore_checkpoint(path, _model()) == (1, 0)
    bad = dict(ck["optimizer"], param_groups=[dict(ck["optimizer"]["param_groups"][0], params=[0, 1, 2])])
    with pytest.raises(ValueError, match="parameters"):
        FusedAdamW(m2).load_state_dict(bad)


# ---------------------------------------------------------------- pinned to the unmodified reference (oracle/make_golden_optim.py)
def _gold():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "optim.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("variant,ncls", [("DFormer-Tiny", 40), ("DFormer-Small", 40), ("DFormer-Base", 37), ("DFormer-Large", 40)])
def test_optimizer_groups_are_the_references_group_weight(variant, ncls):
    """Same parameters, same order (= `torch.optim.AdamW` state numbering) as `group_weight` on the reference model; the fused
    optimizer's per-element decay / update masks agree with the two groups."""
    from dformer_b200.optim import FusedAdamW
    g = _gold()["groups"][variant]
    m = _model(variant)                       # parameter names / groups do not depend on the class count
    opt = FusedAdamW(m, weight_decay=0.01)
    names = {id(p): k for k, p in m.named_parameters()}
    decay, no_decay = opt._reference_groups()
    assert [names[id(p)] for p in decay] == g["decay"] and [names[id(p)] for p in no_decay] == g["no_decay"]
    assert len(names) == g["n_parameters"] and g["no_decay_weight_decay"] == 0.0
    in_decay, in_no_decay = set(g["decay"]), set(g["no_decay"])
    for st in opt.state:
        for s in st["mod"]._plan.layout.slots.values():
            k = names[id(s.param)]
            want = (0.01, 1.0) if k in in_decay else (0.0, 1.0) if k in in_no_decay else (0.0, 0.0)      # neither group: never updated
            assert (float(st["wd"][s.offset]), float(st["lrm"][s.offset])) == pytest.approx(want), k


def test_warmup_poly_lr_matches_the_reference_schedule():
    from dformer_b200.optim import WarmUpPolyLR
    for case in _gold()["lr"]:
        pol = WarmUpPolyLR(case["start_lr"], case["lr_power"], case["total_iters"], case["warmup_steps"])
        for it, want in case["samples"]:
            assert pol.get_lr(it) == want, (case, it)


# ---------------------------------------------------------------- sliding-window bookkeeping (row N3, utils/val_mm.py:287-318)
@pytest.mark.parametrize("h_img,w_img,crop,rate", [(480, 640, (480, 640), 2 / 3), (96, 160, (64, 96), 2 / 3), (100, 100, (64, 64), 0.5),
                                                   (64, 64, (64, 96), 2 / 3), (65, 97, (64, 96), 1.0), (300, 200, (32, 32), 0.75)])
def test_slide_windows_and_counts_match_the_reference_bookkeeping(h_img, w_img, crop, rate):
    from dformer_b200.evaluation import slide_counts, slide_windows
    h_crop, w_crop = crop
    h_stride, w_stride = int(rate * h_crop), int(rate * w_crop)                  # the reference loop, restated
    h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
    count = torch.zeros(h_img, w_img)
    want = []
    for h_idx in range(h_grids):
        for w_idx in range(w_grids):
            y1, x1 = h_idx * h_stride, w_idx * w_stride
            y2, x2 = min(y1 + h_crop, h_img), min(x1 + w_crop, w_img)
            y1, x1 = max(y2 - h_crop, 0), max(x2 - w_crop, 0)
            want.append((y1, y2, x1, x2))
            count[y1:y2, x1:x2] += 1
    wins = slide_windows(h_img, w_img, crop, rate)
    assert wins == want
    rows, cols = slide_counts(h_img, w_img, wins)
    assert torch.equal(torch.tensor(rows, dtype=torch.float32)[:, None] * torch.tensor(cols, dtype=torch.float32)[None, :], count)
    assert count.min() >= 1


# ---------------------------------------------------------------- epoch loop (row N1, utils/train.py:290-470) with fakes
def test_training_loop_follows_the_reference_order_of_operations(tmp_path):
    import os
    from dformer_b200.engine import CheckpointKeeper, is_eval, restore_checkpoint, train
    from dformer_b200.optim import WarmUpPolyLR

    class Opt:                                       # stands in for FusedAdamW (lr, set_lr, state_dict)
        def __init__(self):
            self.lr = 6e-5

        def set_lr(self, lr):
            self.lr = lr

        def state_dict(self):
            return {"lr": self.lr}

        def load_state_dict(self, sd):
            self.lr = sd["lr"]

    class Runner:
        def __init__(self, opt):
            self.opt, self.seen, self.model = opt, [], nn.Linear(2, 2)

        def step(self, rgb, modal_x, label):
            self.seen.append((int(rgb), self.opt.lr))
            return torch.tensor(float(len(self.seen)))

    cfg = SimpleNamespace(nepochs=23, niters_per_epoch=3, checkpoint_start_epoch=20)
    pol = WarmUpPolyLR(6e-5, 0.9, cfg.nepochs * cfg.niters_per_epoch, 2 * cfg.niters_per_epoch)
    opt = Opt()
    run = Runner(opt)
    mious = {1: 0.10, 10: 0.30, 20: 0.25, 21: 0.35, 22: 0.40, 23: 0.45}
    evaluated = []

    def evaluate_fn(epoch):
        evaluated.append(epoch)
        return mious[epoch]

    keeper = CheckpointKeeper(str(tmp_path), keep=3)
    best, hist = train(run, opt, cfg, lambda e: ((torch.tensor(e * 100 + i), None, None) for i in range(10)), pol, evaluate_fn, keeper)
    assert evaluated == [e for e in range(1, 24) if is_eval(e, cfg)] == [1, 10, 20, 21, 22, 23]
    assert best == 0.45 and len(hist) == 23 and hist[0]["miou"] == 0.10 and "miou" not in hist[1]
    # order of operations: the first step runs at the construction-time rate, step k at the rate computed for iteration k - 1
    assert [b for b, _ in run.seen[:4]] == [100, 101, 102, 200]
    want = [6e-5] + [pol.get_lr(k) for k in range(len(run.seen) - 1)]
    assert [lr for _, lr in run.seen] == want
    assert hist[0]["loss"] == pytest.approx(2.0) and hist[-1]["lr"] == pol.get_lr(23 * 3 - 1)
    # new bests at epochs 1, 10, 21, 22, 23 (20 was not a best); the three best files remain
    assert sorted(os.listdir(tmp_path)) == sorted(f"epoch-{e}_miou_{mious[e]}.pt" for e in (21, 22, 23))
    m2, o2 = nn.Linear(2, 2), Opt()
    assert restore_checkpoint(keeper.path(23, 0.45), m2, o2) == (24, 2)                 # epoch + 1, index of the last iteration
    assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), run.model.state_dict().values()))
