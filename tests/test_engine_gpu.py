"""CUDA-graph training step == eager training step (same kernels, captured once, replayed)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _build(seed):
    from dformer_b200 import EncoderDecoder
    cfg = SimpleNamespace(backbone="DFormer-Tiny", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.0, aux_rate=0.0,
                          device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision="bf16", return_logits=False)
    torch.manual_seed(seed)
    m = EncoderDecoder(cfg, norm_layer=nn.BatchNorm2d).cuda().train()
    m.decode_head.dropout = None
    return m


def test_graphed_step_matches_eager_step():
    from dformer_b200.engine import GraphedTrainStep
    from dformer_b200.optim import FusedAdamW
    rgb, hha = torch.randn(2, 3, 96, 128, device="cuda"), torch.randn(2, 3, 96, 128, device="cuda")
    lab = torch.randint(0, 40, (2, 96, 128), device="cuda")
    losses = {}
    finals = {}
    for mode in ("eager", "graph"):
        m = _build(0)
        opt = FusedAdamW(m, lr=1e-3, reference_groups=False)
        torch.manual_seed(1)                              # CPU RNG drives the NMF bases
        run = GraphedTrainStep(m, opt, rgb, hha, lab, warmup=1, use_graph=(mode == "graph"))
        torch.manual_seed(2)
        ls = [run.step(rgb, hha, lab).item() for _ in range(4)]
        losses[mode] = ls
        finals[mode] = m.encoder_backbone.stages[2][0].mlp.fc1.weight.detach().clone()
    # warm-up and capture passes are rolled back (preserve_state), so both runners start from the constructed model and take the
    # same four steps: equal trajectories up to bf16 run-to-run noise (split-K fp32 atomics)
    assert all(torch.isfinite(torch.tensor(v)).all() for v in losses.values())
    assert losses["graph"] == pytest.approx(losses["eager"], abs=2e-2)
    assert (finals["graph"] - finals["eager"]).abs().max() < 6e-3          # 4 AdamW steps at lr 1e-3 move a weight by <= 4e-3
    assert run.graph is not None


def test_graph_replays_advance_the_checkpointed_step_and_eval_draws_fresh_bases(tmp_path):
    """(a) a checkpoint written after N graph replays must carry step == N (the device counter is authoritative) and
    restore into an optimizer that continues with bias corrections of step N + 1; (b) building the runner must not advance the
    model / optimizer state; (c) the runner's injected NMF bases are scoped to its own step, so an evaluation forward with a larger
    batch on the same model draws its own bases instead of reading past the training buffer."""
    from dformer_b200.engine import GraphedTrainStep, restore_checkpoint, save_checkpoint
    from dformer_b200.optim import FusedAdamW
    rgb, hha = torch.randn(2, 3, 64, 96, device="cuda"), torch.randn(2, 3, 64, 96, device="cuda")
    lab = torch.randint(0, 40, (2, 64, 96), device="cuda")
    m = _build(0)
    w0 = m.encoder_backbone.stages[1][0].mlp.fc1.weight.detach().clone()
    rm0 = m.decode_head.squeeze.bn.running_mean.detach().clone()
    opt = FusedAdamW(m, lr=1e-3)
    run = GraphedTrainStep(m, opt, rgb, hha, lab, warmup=2, use_graph=True)
    assert torch.equal(m.encoder_backbone.stages[1][0].mlp.fc1.weight, w0) and torch.equal(m.decode_head.squeeze.bn.running_mean, rm0)
    assert opt.step_count == 0 and float(opt._dyn[1].item()) == 0.0
    N = 5
    for _ in range(N):
        run.step(rgb, hha, lab)
    torch.cuda.synchronize()
    assert float(opt._dyn[1].item()) == N
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, m, opt, epoch=1, iteration=N)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    steps = {int(float(e["step"])) for e in ck["optimizer"]["state"].values()}
    assert steps == {N}
    m2 = _build(1)
    opt2 = FusedAdamW(m2, lr=1.0)
    restore_checkpoint(path, m2, opt2)
    assert opt2.step_count == N
    # (c) evaluation on the trained model with a bigger batch than the runner's
    assert m.decode_head.injected_bases is None
    m.eval()
    with torch.no_grad():
        out = m(torch.randn(3, 3, 64, 96, device="cuda"), torch.randn(3, 3, 64, 96, device="cuda"))
    assert out.shape == (3, 40, 64, 96) and torch.isfinite(out).all()
    m.decode_head.injected_bases = torch.rand(2, 512, 64, device="cuda")
    with pytest.raises(ValueError):
        m(torch.randn(3, 3, 64, 96, device="cuda"), torch.randn(3, 3, 64, 96, device="cuda"))


def test_staged_input_prefetch_feeds_the_same_step():
    """stage() + step() (H2D on the copy stream into staging buffers, consumed by the next step) trains on exactly the batch
    that step(batch) would: with lr = 0 the weights stay put, so equal losses mean equal inputs reached the graph."""
    from dformer_b200.engine import GraphedTrainStep
    from dformer_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(2, 3, 96, 128, generator=g).pin_memory(), torch.randn(2, 3, 96, 128, generator=g).pin_memory(),
                torch.randint(0, 40, (2, 96, 128), generator=g).pin_memory()) for _ in range(3)]
    m = _build(0)
    opt = FusedAdamW(m, lr=0.0, weight_decay=0.0, reference_groups=False)
    b0 = tuple(t.cuda() for t in batches[0])
    run = GraphedTrainStep(m, opt, *b0, warmup=1, use_graph=True)
    direct, staged = [], []
    for b in batches:
        torch.manual_seed(7)                                 # same NMF bases draw for both variants
        direct.append(run.step(*b).item())
    torch.manual_seed(7)                                     # the staged variant draws the bases inside stage()
    run.stage(*batches[0])
    for i in range(len(batches)):
        loss = run.step()
        if i + 1 < len(batches):
            torch.manual_seed(7)
            run.stage(*batches[i + 1])                       # overlaps the step just launched
        staged.append(loss.item())
    # run-to-run noise of a bf16 step is ~1e-4 on the loss (split-K fp32 atomics in the NMF products feed bf16 roundings);
    # the three batches differ by > 2e-3 even at the default init (layer scales 1e-6)
    assert staged == pytest.approx(direct, abs=4e-4)
    assert min(abs(a - b) for i, a in enumerate(direct) for b in direct[i + 1:]) > 1.2e-3


def test_checkpoint_restores_the_trajectory_and_interchanges_with_torch_adamw(tmp_path):
    """Row N1: `save_checkpoint` / `restore_checkpoint` (utils/engine/engine.py:101-186) with the fused optimizer.  After two
    training steps the file is restored (a) into a fresh model + FusedAdamW and (b) into a fresh model driven by the optimizer the
    reference builds -- stock `torch.optim.AdamW` over `group_weight`'s two groups (utils/init_func.py:26-70) -- and all three take
    a third step on the SAME gradients: identical parameters, with `layer_scale_*` / LayerNorm / `stem_e_fc*` untouched."""
    from dformer_b200.engine import restore_checkpoint, save_checkpoint
    from dformer_b200.optim import FusedAdamW
    rgb, hha = torch.randn(2, 3, 64, 96, device="cuda"), torch.randn(2, 3, 64, 96, device="cuda")
    lab = torch.randint(0, 40, (2, 64, 96), device="cuda")
    bases = torch.rand(2, 512, 64, device="cuda")

    def make(seed):
        m = _build(seed)
        m.decode_head.injected_bases = bases
        return m

    def backward(m):
        loss, _ = m(rgb, hha, lab)
        loss.backward()

    m = make(0)
    init = {k: v.detach().clone() for k, v in m.named_parameters()}
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    for _ in range(2):
        backward(m)
        opt.step()
        opt.zero_grad()
    path = str(tmp_path / "epoch-4.pt")
    save_checkpoint(path, m, opt, epoch=4, iteration=2)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"model", "optimizer", "epoch", "iteration"} and list(ck["model"]) == list(m.state_dict())
    # (a) fresh model + fused optimizer
    m2 = make(1)
    opt2 = FusedAdamW(m2, lr=1.0, weight_decay=0.0)
    assert restore_checkpoint(path, m2, opt2) == (5, 2)
    assert (opt2.step_count, opt2.lr, opt2.weight_decay) == (2, 1e-3, 0.05)
    # (b) fresh model + the reference's stock optimizer
    m3 = make(2)
    decay, no_decay = FusedAdamW(m3)._reference_groups()
    opt3 = torch.optim.AdamW([dict(params=decay, lr=1.0), dict(params=no_decay, weight_decay=0.0, lr=1.0)], lr=1.0, weight_decay=0.3)
    assert restore_checkpoint(path, m3, opt3) == (5, 2)
    # third step, same gradients everywhere
    for mm in (m, m2, m3):
        backward(mm)
    torch.cuda.synchronize()
    for src, dst in zip((m.encoder_backbone, m.decode_head), (m2.encoder_backbone, m2.decode_head)):
        dst._last_arena.buf.copy_(src._last_arena.buf)                # the fused optimizer reads the gradient arena
    g1 = {k: p.grad for k, p in m.named_parameters()}
    for k, p in m3.named_parameters():
        p.grad = None if g1[k] is None else g1[k].detach().clone()    # the stock optimizer reads p.grad
    opt.step()
    opt2.step()
    opt3.step()
    torch.cuda.synchronize()
    p1, p2, p3 = (dict(mm.named_parameters()) for mm in (m, m2, m3))
    moved = 0
    for k in p1:
        torch.testing.assert_close(p2[k], p1[k], rtol=0, atol=0, msg=lambda s, k=k: f"{k}: {s}")
        torch.testing.assert_close(p3[k], p1[k], rtol=1e-5, atol=1e-6, msg=lambda s, k=k: f"{k}: {s}")
        frozen = "layer_scale" in k or ".norm" in k or "stem_e_fc" in k
        if frozen:
            assert torch.equal(p1[k], init[k]), k                      # the reference's optimizer never sees these
        else:
            moved += int(not torch.equal(p1[k], init[k]))
    assert moved > 300
