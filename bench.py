"""bench.py -- DFormer-L 480x640 bf16 TRAINING throughput (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                      (single GPU)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference [...]                              (reference CPU path = oracle port)

A step = forward + loss + backward (+ NCCL gradient all-reduce when N > 1) + fused AdamW of
DFormer-Large + LightHamHead on a synthetic 480x640 RGB+HHA batch of 8 images per GPU (BASELINE.json
configs[3]: global batch 64 on 8 GPUs; weak scaling).  `value` is timed with inputs resident in HBM;
`e2e` times the same step through the public API with pinned-host inputs copied H2D and the loss read back
D2H inside the timed region.  Random-init weights, synthetic data (no datasets / checkpoints offline).

Beside the headline the line carries: `roofline` (dominant kernel = the tcgen05 GEMM family, every launch configuration of the
step timed live over rotating operand buffers larger than the L2), `other_kernels` (worst shape per family), `cpu_baseline`
(oracle port on the host cores), `gpu_eager_reference` (the reference's ops as stock PyTorch eager on the same B200),
`dp_check` (N > 1: gradients of the N-rank exchange vs one rank on the concatenated batch) and `extra` (the other BASELINE.json
configurations: Small / Base training, the DFormer-L inference sweep, the Tiny fp32 CPU forward)."""
import argparse
import json
import os
import sys
import threading
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_IMG_TRAIN = 411.6e9          # 3 x train-mode forward of DFormer-L @480x640/40cls (SURVEY.md 8d / BASELINE.md)
VARIANT, NCLS, H, W, PER_GPU_BATCH = "DFormer-Large", 40, 480, 640, 8
METRIC = "DFormer-L 480x640 bf16 train images/s"
L2_BYTES = 126 << 20


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region (pynvml; nvidia-smi equivalent)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("sampler_error:" + type(e).__name__)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def synthetic(batch, seed, ncls=None, h=None, w=None):
    ncls, h, w = ncls or NCLS, h or H, w or W
    g = torch.Generator(device="cpu").manual_seed(seed)
    rgb = torch.randn(batch, 3, h, w, generator=g)
    hha = torch.randn(batch, 3, h, w, generator=g)
    label = torch.randint(0, ncls, (batch, h, w), generator=g)
    label[torch.rand(batch, h, w, generator=g) < 0.05] = 255
    return rgb, hha, label


def cfg_for(precision, device, variant=None, ncls=None):
    # drop_path of the matching local_configs file: 0.15 for Large, 0.1 otherwise
    variant, ncls = variant or VARIANT, ncls or NCLS
    return SimpleNamespace(backbone=variant, decoder="ham", decoder_embed_dim=512, num_classes=ncls,
                           drop_path_rate=0.15 if variant == "DFormer-Large" else 0.1, aux_rate=0.0,
                           device=device, pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision=precision)


# ------------------------------------------------------------------------------------------ reference arms (oracle; no product code)
def _oracle_state(variant, ncls, device="cpu"):
    """A random state_dict in the reference's key layout WITHOUT touching the product package (so that the reference arm's process
    never maps libdformer_b200.so): shapes from the committed layout fixture (generated from the unmodified reference by
    oracle/make_golden.py), values from the fixture generator, layer scales at the reference's 1e-6 (DFormer.py:152)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import make_state
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_layouts.json")) as f:
        lay = json.load(f)[variant]
    shapes = {k: tuple(s) for k, s in lay["shapes"].items()}
    for k in list(shapes):
        if k.startswith("decode_head.conv_seg"):
            shapes[k] = (ncls,) + shapes[k][1:]
    P = make_state(shapes, seed=0)
    for k in P:
        if ".layer_scale_" in k:
            P[k].fill_(1e-6)
    P = {k: v.to(device) for k, v in P.items()}
    train = [k for k in lay["trainable"] if not k.startswith("encoder_backbone.stem_e_fc")]
    for k in train:
        P[k].requires_grad_(True)
    return P, train


def cpu_reference_run(steps, warmup, batch=2):
    """The reference's own CPU implementation of the path = the oracle port in its ATen-faithful mode (F.layer_norm / F.gelu /
    F.batch_norm / ... as the reference's modules dispatch them), fp32, all host cores, torch.optim.AdamW like utils/train.py:211.
    A step here is a bounded sample of the headline step: `batch` of its 8 images."""
    from oracle import dformer_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P, train = _oracle_state(VARIANT, NCLS)
    opt = torch.optim.AdamW([P[k] for k in train], lr=6e-5, weight_decay=0.01)
    rgb, hha, label = synthetic(batch, 0)
    v = O.VARIANTS[VARIANT]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        bases = O.draw_bases(batch)
        with O.aten_faithful():
            loss, _ = O.forward(P, rgb, hha, bases, v["dims"], v["depths"], label=label, training=True)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec * 1e3, torch.get_num_threads(), (f"{steps} train steps (fwd+loss+bwd+AdamW, fp32, ATen ops) of {VARIANT} {H}x{W} on "
                                                             f"{batch} of the step's {PER_GPU_BATCH} images, after {warmup} warm-up")


def cpu_tiny_forward(reps=5):
    """BASELINE.json configs[0]: DFormer-Tiny + LightHamHead fp32 eval forward, batch 1, 480x640, 40 classes, on the host cores."""
    from oracle import dformer_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    P, _ = _oracle_state("DFormer-Tiny", 40)
    rgb, hha, _ = synthetic(1, 0, 40, 480, 640)
    v = O.VARIANTS["DFormer-Tiny"]
    ts = []
    with torch.no_grad(), O.aten_faithful():
        for it in range(reps + 2):
            t0 = time.perf_counter()
            O.forward(P, rgb, hha, O.draw_bases(1), v["dims"], v["depths"], training=False)
            if it >= 2:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return {"workload": "DFormer-Tiny + LightHamHead fp32 eval forward, batch 1, 480x640, 40 cls, CPU (oracle port, ATen ops)",
            "median_ms": ts[len(ts) // 2] * 1e3, "cores": torch.get_num_threads(), "reps": reps}


def gpu_eager_reference(steps, warmup, batch, precision, variant=None, ncls=None, h=None, w=None):
    """SURVEY.md 8(d) "the real competitor": the reference's algorithm as STOCK PyTorch eager ops on the same B200 -- the
    ATen-faithful oracle (the ops the reference's nn.Modules dispatch) under autograd, `torch.autocast(bf16)` for the bf16 row,
    `torch.optim.AdamW` for the update.  Not product code, not credited: a comparator."""
    from oracle import dformer_oracle as O
    variant, ncls, h, w = variant or VARIANT, ncls or NCLS, h or H, w or W
    dev = torch.device("cuda", torch.cuda.current_device())
    P, train = _oracle_state(variant, ncls, dev)
    opt = torch.optim.AdamW([P[k] for k in train], lr=6e-5, weight_decay=0.01)
    rgb, hha, label = (t.to(dev) for t in synthetic(batch, 0, ncls, h, w))
    v = O.VARIANTS[variant]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    for it in range(warmup + steps):
        if it == warmup:
            torch.cuda.synchronize()
            e0.record()
        bases = O.draw_bases(batch).to(dev, non_blocking=True)
        with O.aten_faithful(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(precision == "bf16")):
            loss, _ = O.forward(P, rgb, hha, bases, v["dims"], v["depths"], label=label, training=True)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
            "what": f"ATen-faithful oracle port, stock PyTorch eager on this GPU, {'torch.autocast(bf16)' if precision == 'bf16' else 'fp32'}, "
                    f"torch.optim.AdamW, {variant} {h}x{w} batch {batch}, drop_path 0, {steps} steps after {warmup} warm-up"}


# ------------------------------------------------------------------------------------------ product-path helpers
def _time_replays(fn, reps, stream):
    """device time of one call of fn (a sequence of launches): CUDA events around a captured graph of `reps` calls"""
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for i in range(reps):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    g.replay()
    q1.record()
    torch.cuda.synchronize()
    return q0.elapsed_time(q1) / reps * 1e-3


def train_config_ms(variant, h, w, ncls, batch, steps, warmup, precision="bf16", world=1, dev=None):
    """graph-replayed training step of another BASELINE.json configuration (same engine, same code path as the headline)"""
    import torch.distributed as dist
    import torch.nn as nn
    from dformer_b200 import EncoderDecoder
    from dformer_b200.engine import GraphedTrainStep
    from dformer_b200.optim import FusedAdamW
    from dformer_b200.parallel import GradSync
    torch.manual_seed(0)
    model = EncoderDecoder(cfg_for(precision, "cuda", variant, ncls), norm_layer=nn.SyncBatchNorm if world > 1 else nn.BatchNorm2d,
                           syncbn=(world > 1)).to(dev).train()
    model.cfg.return_logits = False
    opt = FusedAdamW(model, lr=6e-5, weight_decay=0.01)
    sync = GradSync(model)
    rgb, hha, lab = (t.to(dev) for t in synthetic(batch, 7, ncls, h, w))
    runner = GraphedTrainStep(model, opt, rgb, hha, lab, grad_sync=sync, warmup=2, use_graph=True)
    for _ in range(warmup):
        runner.step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        runner.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    del runner, model, opt, sync
    torch.cuda.empty_cache()
    return {"workload": f"{variant} + LightHamHead {h}x{w} {ncls}cls {precision} train step (fwd+loss+bwd+AdamW), batch {batch}/GPU x {world} GPU, cuda graph",
            "ms_per_step": ms, "images_per_s": world * batch / (ms * 1e-3), "steps": steps, "warmup": warmup}


def inference_sweep(dev, batches=(1, 8, 64), precisions=("bf16", "fp32")):
    """BASELINE.json configs[4]: DFormer-L eval forward at 480x640, graph-replayed (protocol of utils/latency.py:27-63: CUDA events
    after warm-up), bf16 and fp32 (the exact CUDA-core fp32 path: tcgen05 has no fp32 MMA)."""
    import torch.nn as nn
    from dformer_b200 import EncoderDecoder
    rows = []
    for prec in precisions:
        torch.manual_seed(0)
        m = EncoderDecoder(cfg_for(prec, "cuda"), norm_layer=nn.BatchNorm2d).to(dev).eval()
        for B in (batches if prec == "bf16" else batches[:1]):
            rgb, hha = torch.rand(B, 3, H, W, device=dev), torch.rand(B, 3, H, W, device=dev)
            m.decode_head.injected_bases = torch.rand(B, 512, 64, device=dev)
            with torch.no_grad():
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(2):
                        m(rgb, hha)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = m(rgb, hha)
                reps = 50 if B == 1 else 10
                for _ in range(5):
                    g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rows.append({"batch": B, "precision": prec, "graph_ms": ms, "images_per_s": B / ms * 1e3})
            del g, out
        m.decode_head.injected_bases = None
        del m
        torch.cuda.empty_cache()
    return rows


def dp_equivalence(rank, world, dev):
    """N-rank gradients after the NCCL exchange (SyncBN on) vs ONE rank on the concatenated batch (plain BN), DropPath / dropout
    off, fixed NMF bases: DFormer-Tiny 64x96, fp32 (exact CUDA-core path) and bf16.  Every rank runs the sharded step; rank 0
    additionally runs the whole batch alone and reports the minimum per-tensor gradient cosine."""
    import torch.nn as nn
    from dformer_b200 import EncoderDecoder
    from dformer_b200.parallel import GradSync
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import make_inputs, make_state
    per = 2
    rgb, hha, label, bases = make_inputs(per * world, 64, 96, 40, seed=9)
    label[label == 255] = 0                                      # equal valid-pixel counts per shard (the loss is a local mean)
    sl = slice(rank * per, (rank + 1) * per)

    def build(precision, syncbn):
        cfg = SimpleNamespace(backbone="DFormer-Tiny", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.0, aux_rate=0.0,
                              device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision=precision)
        m = EncoderDecoder(cfg, norm_layer=nn.SyncBatchNorm if syncbn else nn.BatchNorm2d, syncbn=syncbn)
        m.load_state_dict(make_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=3))
        m.to(dev).train()
        m.decode_head.dropout = None
        # the four stem BatchNorms stay per-rank nn.BatchNorm2d in the reference even under SyncBN (DFormer.py:196-210): frozen here
        for seq in (m.encoder_backbone.downsample_layers[0], m.encoder_backbone.downsample_layers_e[0]):
            seq[1].eval()
            seq[4].eval()
        return m

    res = {"world": world, "model": "DFormer-Tiny 64x96, 2 images per rank"}
    for precision in ("fp32", "bf16"):
        m = build(precision, True)
        sync = GradSync(m, bucket_mb=1.0)
        m.decode_head.injected_bases = bases[sl].to(dev)
        loss, _ = m(rgb[sl].to(dev), hha[sl].to(dev), label[sl].to(dev))
        loss.backward()
        sync.finish()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
        if rank == 0:
            ref = build(precision, False)
            ref.decode_head.injected_bases = bases.to(dev)
            l2, _ = ref(rgb.to(dev), hha.to(dev), label.to(dev))
            l2.backward()
            worst = 1.0
            for k, p in ref.named_parameters():
                if p.grad is None or p.grad.norm() < (1e-6 if precision == "fp32" else 1e-3):    # numerically-zero gradients: noise only
                    continue
                worst = min(worst, torch.nn.functional.cosine_similarity(p.grad.flatten().float(), grads[k].flatten().float(), dim=0).item())
            res[precision] = {"min_grad_cosine_vs_1_rank": worst, "buckets_all_reduced": sync.launched}
            del ref
        del m, sync
    if rank == 0:
        res["ok"] = bool(res["fp32"]["min_grad_cosine_vs_1_rank"] >= 0.999999 and res["bf16"]["min_grad_cosine_vs_1_rank"] >= 0.99)
    torch.cuda.empty_cache()
    return res


def gemm_census_roofline(runner, dev, K):
    """Dominant kernel: the tcgen05 GEMM family.  Census of one step (shape, algorithmic bytes / FLOPs, count), then every distinct
    launch configuration is timed live: CUDA events around a captured graph of back-to-back launches that ROTATE through as many
    operand / output buffer sets as it takes to exceed twice the L2 (so no launch finds its operands cached by the previous one)."""
    from collections import Counter
    K.GEMM_PROFILE, K.GEMM_SHAPES_ONLY = [], True
    runner._draw_bases()
    runner._eager()
    torch.cuda.synchronize()
    census = Counter((fl, by) + shp for _, _, fl, by, is_tc, shp in K.GEMM_PROFILE if is_tc)
    n_simt = sum(1 for r in K.GEMM_PROFILE if not r[4])
    K.GEMM_PROFILE, K.GEMM_SHAPES_ONLY = None, False
    tot = dict(ms=0.0, fl=0.0, by=0.0, n=0)
    cap_stream = torch.cuda.Stream()
    worst = None
    for (fl, by, M_, N_, K_, ta, tb, f32out, has_bias, acc), cnt in census.items():
        per_set = (M_ * K_ + N_ * K_) * 2 + M_ * N_ * (4 if f32out else 2)
        nsets = max(2, min(24, -(-2 * L2_BYTES // per_set)))
        sets = []
        for _ in range(nsets):
            a_ = (torch.randn(K_, M_, device=dev) if ta else torch.randn(M_, K_, device=dev)).bfloat16()
            b_ = (torch.randn(N_, K_, device=dev) if tb else torch.randn(K_, N_, device=dev)).bfloat16()
            o_ = torch.zeros(M_, N_, device=dev, dtype=torch.float32 if f32out else torch.bfloat16)
            sets.append((a_, b_, o_))
        bias_ = torch.zeros(N_, device=dev) if has_bias else None

        def launch(i, sets=sets, ta=ta, tb=tb, acc=acc, bias_=bias_):
            a_, b_, o_ = sets[i % len(sets)]
            K.gemm(a_, b_, trans_a=bool(ta), trans_b=bool(tb), backend=K.TCGEN05, out=o_, accumulate=bool(acc), bias=bias_)

        reps = max(20, 2 * nsets)
        for i in range(nsets):
            launch(i)
        per = _time_replays(launch, reps, cap_stream)
        K.gemm_sm_budget(148, 148)                          # the same launches with grids over the whole chip (the kernel alone)
        launch(0)
        tot["ms_full_chip"] = tot.get("ms_full_chip", 0.0) + _time_replays(launch, reps, cap_stream) * 1e3 * cnt
        K.gemm_sm_budget(0, 0)
        tot["ms"] += per * 1e3 * cnt
        tot["fl"] += fl * cnt
        tot["by"] += by * cnt
        tot["n"] += cnt
        gbs = by / per / 1e9
        if cnt >= 3 and (worst is None or gbs < worst["achieved"]):
            worst = {"shape_MNK": [M_, N_, K_], "transA": int(ta), "transB": int(tb), "launches_per_step": cnt, "us": per * 1e6,
                     "achieved": gbs, "unit": "GB/s", "tflops": fl / per / 1e12}
        del sets
    return tot, n_simt, worst


def other_kernel_rooflines(dev, K, B, hbm):
    """HBM / FMA-bound kernels next in line, timed stand-alone over rotating buffers; the WORST shape of each family in the step is
    the headline entry (`frac`), the best is listed beside it."""
    cap_stream = torch.cuda.Stream()
    rows = []
    ffma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    try:
        fam = {"mlp_dw_fwd_kernel (dw3x3 + residual + GELU, keeps GELU')": [], "mlp_dw_bwd_saved_kernel (dz, dh, dW, db, fc1 bias grad)": [],
               "dw7_conv_kernel (depthwise 7x7)": []}
        for (hs, ws, ch, c7) in ((H // 4, W // 4, 768, 96), (H // 8, W // 8, 1536, 192), (H // 16, W // 16, 1152, 288), (H // 32, W // 32, 2304, 576)):
            M_ = B * hs * ws
            nsets = max(2, min(8, -(-2 * L2_BYTES // (M_ * ch * 2 * 3))))
            hh = [torch.randn(M_, ch, device=dev).bfloat16() for _ in range(nsets)]
            du = [torch.randn_like(t) for t in hh]
            w3, b3 = torch.randn(ch, 1, 3, 3, device=dev) * 0.2, torch.randn(ch, device=dev) * 0.1
            dw3, db3, dc3 = torch.zeros_like(w3), torch.zeros_like(b3), torch.zeros(ch, device=dev)
            gp = [K.mlp_dw_fwd(t, w3, b3, B, hs, ws, save_gp=True)[1] for t in hh]
            by = M_ * ch * 2
            t_f = _time_replays(lambda i: K.mlp_dw_fwd(hh[i % nsets], w3, b3, B, hs, ws, save_gp=True), 2 * nsets, cap_stream)
            t_b = _time_replays(lambda i: K.mlp_dw_bwd(du[i % nsets], hh[i % nsets], w3, b3, B, hs, ws, dw3, db3, dc3, gp=gp[i % nsets]), 2 * nsets, cap_stream)
            fam["mlp_dw_fwd_kernel (dw3x3 + residual + GELU, keeps GELU')"].append(dict(shape=[B, hs, ws, ch], algorithmic_bytes=3 * by, us=t_f * 1e6, achieved=3 * by / t_f / 1e9))
            fam["mlp_dw_bwd_saved_kernel (dz, dh, dW, db, fc1 bias grad)"].append(dict(shape=[B, hs, ws, ch], algorithmic_bytes=4 * by, us=t_b * 1e6, achieved=4 * by / t_b / 1e9))
            del hh, du, gp
            n7 = max(2, min(8, -(-2 * L2_BYTES // (M_ * c7 * 4))))
            x7 = [torch.randn(M_, c7, device=dev).bfloat16() for _ in range(n7)]
            w7, b7 = torch.randn(c7, 1, 7, 7, device=dev) * 0.1, torch.randn(c7, device=dev) * 0.1
            t_7 = _time_replays(lambda i: K.dwconv_fwd(x7[i % n7], w7, b7, B, hs, ws, 7), 2 * n7, cap_stream)
            fl7 = 2.0 * 49 * M_ * c7
            fam["dw7_conv_kernel (depthwise 7x7)"].append(dict(shape=[B, hs, ws, c7], algorithmic_flops=fl7, us=t_7 * 1e6, achieved=fl7 / t_7 / 1e12))
            del x7
        for name, lst in fam.items():
            fma = "dw7" in name
            lst.sort(key=lambda r: r["achieved"])
            peak = ffma_peak if fma else hbm
            rows.append(dict(lst[0], kernel=name, bound="fp32 FMA" if fma else "hbm", unit="TFLOP/s" if fma else "GB/s", peak=peak,
                             frac=lst[0]["achieved"] / peak, which="worst shape of the step's four stages",
                             best=dict(shape=lst[-1]["shape"], achieved=lst[-1]["achieved"], frac=lst[-1]["achieved"] / peak)))
    except Exception as e:  # noqa: BLE001
        rows.append({"error": f"{type(e).__name__}: {e}"})
    return rows


def main():
    global VARIANT, NCLS, H, W
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the side configurations / eager comparator / dp_check objects")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--variant", default=VARIANT, help="other BASELINE.json configs (use with --quick): DFormer-Tiny/Small/Base/Large")
    ap.add_argument("--size", default=f"{H}x{W}", help="HxW of the synthetic batch (480x480 for the SUNRGBD-shaped config)")
    ap.add_argument("--classes", type=int, default=NCLS)
    ap.add_argument("--torch-eager-gpu", action="store_true", help="context line (not a bench arm): the ATen-faithful oracle port as stock "
                    "PyTorch eager ops on cuda:0, bf16-autocast and fp32, same workload")
    ap.add_argument("--quick", action="store_true", help="profiling aid: print only the device-timed ms/step and exit (not a bench line)")
    args = ap.parse_args()
    custom = (args.variant, args.size, args.classes) != (VARIANT, f"{H}x{W}", NCLS)
    VARIANT, NCLS = args.variant, args.classes
    H, W = (int(t) for t in args.size.split("x"))
    if custom and not (args.quick or args.torch_eager_gpu):
        ap.error("--variant/--size/--classes select a side configuration: combine them with --quick (ms/step only); the bench line "
                 "is defined on the headline configuration")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.torch_eager_gpu:
        if rank != 0:
            return
        rows = {}
        for prec in ("bf16", "fp32"):
            try:
                rows[prec] = gpu_eager_reference(args.steps, args.warmup, args.batch, prec)
            except Exception as e:  # noqa: BLE001  (e.g. out of memory for the fp32 row at a large batch)
                rows[prec] = {"value": None, "unit": "images/s", "ms_per_step": None, "error": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.empty_cache()
        print(json.dumps({"impl": "reference-port (ATen-faithful), stock PyTorch eager on cuda:0", "metric": METRIC, "value": rows["bf16"]["value"],
                          "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rows["bf16"]["ms_per_step"],
                          "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": rows["bf16"].get("what")}, "bf16_autocast": rows["bf16"], "fp32_torch_defaults": rows["fp32"]}))
        return
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 1))        # ~2.5 s per 2-image step on 16 cores: bounded
        val, ms, cores, sample = cpu_reference_run(steps, warmup)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
                          "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"{VARIANT} + LightHamHead {H}x{W} {NCLS}cls train step (fwd+loss+bwd+AdamW), CPU fp32, "
                                                 f"2 of the step's {PER_GPU_BATCH} images per timed step, all host cores"},
                          "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    import torch.nn as nn
    from dformer_b200 import EncoderDecoder, kernels as K
    from dformer_b200._lib import lib
    from dformer_b200.engine import GraphedTrainStep
    from dformer_b200.optim import FusedAdamW
    from dformer_b200.parallel import GradSync

    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dp = None
    if world > 1 and not (args.quick or args.no_extra):
        try:
            dp = dp_equivalence(rank, world, dev)
        except Exception as e:  # noqa: BLE001
            dp = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    torch.manual_seed(0)
    model = EncoderDecoder(cfg_for(args.precision, "cuda"), norm_layer=nn.SyncBatchNorm if world > 1 else nn.BatchNorm2d,
                           syncbn=(world > 1)).to(dev).train()
    model.cfg.return_logits = False        # training consumes the loss only; the fused upsample+CE never materialises 98 MB/img of logits
    opt = FusedAdamW(model, lr=6e-5, weight_decay=0.01)
    sync = GradSync(model)
    B = args.batch
    rgb_h, hha_h, lab_h = (t.pin_memory() for t in synthetic(B, 100 + rank))
    rgb, hha, lab = rgb_h.to(dev), hha_h.to(dev), lab_h.to(dev)
    launches0 = lib().launch_count()
    try:
        runner = GraphedTrainStep(model, opt, rgb, hha, lab, grad_sync=sync, warmup=2, use_graph=not args.no_graph)
        graphed = runner.graph is not None
    except Exception as e:  # noqa: BLE001  (capture is an optimisation; the eager path is the same code)
        print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); falling back to eager launches", file=sys.stderr)
        runner = GraphedTrainStep(model, opt, rgb, hha, lab, grad_sync=sync, warmup=1, use_graph=False)
        graphed = False
    launches_per_step = (lib().launch_count() - launches0) // (3 if graphed else 2) if graphed else None

    def step(r=None, h=None, l=None):
        return runner.step(r, h, l)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib().launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = (launches_per_step * args.steps) if graphed else (lib().launch_count() - launches0)
    ms = e0.elapsed_time(e1) / args.steps
    if args.quick:
        sampler.stop_flag = True
        if rank == 0:
            print(json.dumps({"quick": True, "workload": f"{VARIANT} {H}x{W} {NCLS}cls batch {B}/GPU x {world} GPU, {args.precision}",
                              "ms_per_step": ms, "images_per_s": world * B / (ms * 1e-3), "skip": os.environ.get("DFB200_PROFILE_SKIP", "")}))
        sys.stdout.flush()
        os._exit(0)
    # ---- end-to-end: pinned host inputs -> H2D, step, loss -> D2H, every step.
    # Each step's batch goes pinned host -> device through runner.stage(): a copy stream fills staging buffers while the previous
    # step is still running (the double-buffered prefetch of any training input pipeline), the step consumes them, and its loss
    # comes back device -> host with .item() before the next step is issued.  Everything is inside the timed region.
    runner.stage(rgb_h, hha_h, lab_h)                      # untimed: creates the copy stream / staging / pinned buffers once
    step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    host_loss = 0.0
    runner.stage(rgb_h, hha_h, lab_h)
    for i in range(args.steps):
        loss_dev = step()                                   # consumes the staged batch
        if i + 1 < args.steps:
            runner.stage(rgb_h, hha_h, lab_h)               # H2D of the next step's inputs, overlapping this step
        host_loss = loss_dev.item()                         # D2H of this step's loss (synchronises)
    t1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_e2e = t0.elapsed_time(t1) / args.steps
    tot, n_simt, worst_gemm = gemm_census_roofline(runner, dev, K)
    hbm, tf_burst, tf_sus, how = peaks()
    other = other_kernel_rooflines(dev, K, B, hbm)

    if world > 1:
        tms = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tms.tolist()

    # ---- side configurations (BASELINE.json configs 1, 2, 3, 5) and the eager comparator; N > 1 runs config 3 (Base, data-parallel)
    extra = {}
    del runner
    torch.cuda.empty_cache()
    if not args.no_extra:
        try:
            if world == 1:
                extra["small_480x640_train"] = train_config_ms("DFormer-Small", 480, 640, 40, 8, 5, 3, dev=dev)
                extra["base_480x480_train"] = train_config_ms("DFormer-Base", 480, 480, 37, 8, 5, 3, dev=dev)
                extra["large_inference_480x640"] = inference_sweep(dev)
            else:
                extra["base_480x480_train_dp"] = train_config_ms("DFormer-Base", 480, 480, 37, 8, 5, 3, world=world, dev=dev)
        except Exception as e:  # noqa: BLE001
            extra["error"] = f"{type(e).__name__}: {str(e)[:300]}"

    def leave():
        # NCCL communicators captured inside a CUDA graph can stall a graceful teardown: flush and leave directly
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        leave()
        return
    n = world
    value = n * B / (ms * 1e-3)
    e2e = n * B / (ms_e2e * 1e-3)
    h2d = rgb_h.numel() * 4 + hha_h.numel() * 4 + lab_h.numel() * 8
    tc_ms, tc_fl, tc_by, tc_n = tot["ms"], tot["fl"], tot["by"], tot["n"]
    achieved = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    achieved_gbs = tc_by / (tc_ms * 1e-3) / 1e9 if tc_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:            # DRAM bytes per launch of the same kernel from the committed ncu capture of this round (profiles/); static, not re-measured here
        with open(os.path.join(ROOT, "profiles", "ncu_gemm_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["dram_bytes_per_launch"], "static: " + tj.get("source", "profiles/ncu_gemm_traffic.json")
    except Exception:  # noqa: BLE001
        pass
    out = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": f"{VARIANT} + LightHamHead {H}x{W} {NCLS}cls train step (fwd+loss+bwd+AdamW), batch {B}/GPU, drop_path 0.15, "
                               f"{'SyncBN + NCCL grad all-reduce' if n > 1 else 'single GPU'}; the step returns the loss only "
                               "(cfg.return_logits=False: utils/train.py ignores `out` of builder.py:233's (loss, out))",
                   "global_batch": n * B, "parallelism": f"dp{n}", "launch": "cuda_graph" if graphed else "eager",
                   "l2_policy": f"per-step working set ({B} x ~1.3 GB activations) exceeds the 126 MB L2; no flush needed; the stand-alone "
                                "kernel timings rotate through buffer sets of > 2x the L2"},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "loss": host_loss},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        # DFormer's GEMMs have K = 48..288 for most layers (arithmetic intensity ~ K*N/(K+N) < the 246 FLOP/B ridge), so the
        # dominant kernel is bounded by HBM, not by the tensor pipe; both views are reported.
        "roofline": {"bound": "hbm", "kernel": "gemm_tc_kernel (tcgen05 GEMM: every Linear / 1x1 / im2col conv fwd + dgrad + wgrad)",
                     "achieved": achieved_gbs, "peak": hbm, "unit": "GB/s", "frac": achieved_gbs / hbm, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)",
                     "grid_policy": "persistent grids cover 124 SMs (forward / dgrad) and 72 SMs (split-K weight gradients) of 148: fastest "
                                    "inside the four-stream step (profiles/r02_sm_budget_sweep.txt); `full_chip` = the same launches with grids "
                                    "over all 148 SMs, i.e. the kernel running alone",
                     "full_chip": {"achieved": tc_by / (tot["ms_full_chip"] * 1e-3) / 1e9, "frac": tc_by / (tot["ms_full_chip"] * 1e-3) / 1e9 / hbm,
                                   "kernel_ms_per_step": tot["ms_full_chip"]} if tot.get("ms_full_chip") else None,
                     "algorithmic_bytes_per_launch": tc_by / max(tc_n, 1), "launches_per_step": tc_n,
                     "avg_launch_us": tc_ms * 1e3 / max(tc_n, 1), "kernel_ms_per_step": tc_ms, "kernel_share_of_step": tc_ms / ms,
                     "timing": "per distinct launch configuration of the step: CUDA events around a captured graph of back-to-back launches "
                               "rotating through operand/output buffer sets of > 2x the 126 MB L2 (cold operands), weighted by the "
                               "configuration's launch count in one training step",
                     "tensor_view": {"achieved_tflops": achieved, "peak_tflops": tf_burst, "frac": achieved / tf_burst,
                                     "algorithmic_flops_per_step": tc_fl},
                     "worst_shape": worst_gemm,
                     "cuda_core_gemm_launches_per_step": int(n_simt)},
        "other_kernels": other,
        "step_roofline": {"achieved_tflops": value * FLOP_PER_IMG_TRAIN / n / 1e12, "frac_of_burst": value * FLOP_PER_IMG_TRAIN / n / 1e12 / tf_burst,
                          "frac_of_sustained": value * FLOP_PER_IMG_TRAIN / n / 1e12 / tf_sus, "flop_per_image": FLOP_PER_IMG_TRAIN},
    }
    if dp is not None:
        out["dp_check"] = dp
    if n == 1 and not args.no_extra:
        try:
            out["gpu_eager_reference"] = gpu_eager_reference(3, 2, B, "bf16")
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out["gpu_eager_reference"] = {"value": None, "error": f"{type(e).__name__}: {str(e)[:200]}"}
    if n == 1 and not args.no_cpu_baseline:
        try:
            val, cms, cores, sample = cpu_reference_run(steps=4, warmup=1)          # ~15 s of CPU work
            out["cpu_baseline"] = {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample, "ms_per_step": cms}
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": f"failed: {e}"}
        if not args.no_extra:
            try:
                extra["tiny_cpu_forward"] = cpu_tiny_forward()
            except Exception as e:  # noqa: BLE001
                extra["tiny_cpu_forward"] = {"error": str(e)[:200]}
    if extra:
        out["extra"] = extra
    print(json.dumps(out))
    leave()


if __name__ == "__main__":
    main()
