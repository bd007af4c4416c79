"""DP-equivalence check (run under torchrun, >= 2 GPUs): N-rank gradients after the NCCL exchange == 1-rank
gradients on the concatenated batch (SyncBN on, DropPath/dropout off, fixed NMF bases).  Exits non-zero on mismatch."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from golden_util import make_inputs, make_state  # noqa: E402

from dformer_b200 import EncoderDecoder  # noqa: E402
from dformer_b200.parallel import GradSync  # noqa: E402


def build(precision, syncbn):
    cfg = SimpleNamespace(backbone="DFormer-Tiny", decoder="ham", decoder_embed_dim=512, num_classes=40, drop_path_rate=0.0, aux_rate=0.0,
                          device="cuda", pretrained_model=None, bn_eps=1e-3, bn_momentum=0.1, background=255, precision=precision)
    # utils/train.py:182-194: with --syncbn the reference passes norm_layer = nn.SyncBatchNorm (so init_weight sets the head's BN eps)
    m = EncoderDecoder(cfg, norm_layer=nn.SyncBatchNorm if syncbn else nn.BatchNorm2d, syncbn=syncbn)
    m.load_state_dict(make_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=3))
    m.cuda().train()
    m.decode_head.dropout = None
    # The reference keeps the four stem BatchNorms as plain nn.BatchNorm2d even under SyncBN (DFormer.py:196-210), so with
    # batch statistics they are per-rank by design; freeze them (running stats) to compare N ranks against one exactly.
    for seq in (m.encoder_backbone.downsample_layers[0], m.encoder_backbone.downsample_layers_e[0]):
        seq[1].eval()
        seq[4].eval()
    return m


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per = 2
    rgb, hha, label, bases = make_inputs(per * world, 64, 96, 40, seed=9)
    label = label.clamp(max=39)                                  # no ignored pixels: equal valid counts per shard
    label[label == 255] = 0
    sl = slice(rank * per, (rank + 1) * per)
    ok = True
    for precision, tol in (("fp32", 1e-4), ("bf16", 3e-2)):
        m = build(precision, syncbn=True)
        sync = GradSync(m, bucket_mb=1.0)
        m.decode_head.injected_bases = bases[sl].cuda()
        loss, _ = m(rgb[sl].cuda(), hha[sl].cuda(), label[sl].cuda())
        loss.backward()
        sync.finish()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
        stats = {k: v.detach().clone() for k, v in m.state_dict().items() if k.endswith("running_var")}
        if rank == 0:
            ref = build(precision, syncbn=False)                 # one process, whole batch, plain BN == SyncBN over shards
            ref.decode_head.injected_bases = bases.cuda()
            l2, _ = ref(rgb.cuda(), hha.cuda(), label.cuda())
            l2.backward()
            worst, worst_k = 1.0, None
            for k, p in ref.named_parameters():
                if p.grad is None or p.grad.norm() < (1e-6 if precision == "fp32" else 1e-3):   # (numerically) zero gradients: noise only
                    continue
                cos = torch.nn.functional.cosine_similarity(p.grad.flatten().float(), grads[k].flatten().float(), dim=0).item()
                if cos < worst:
                    worst, worst_k = cos, k
            print(f"[dp_check] worst parameter: {worst_k} |ref grad| {dict(ref.named_parameters())[worst_k].grad.norm().item():.3e}", flush=True)
            rs = {k: v for k, v in ref.state_dict().items() if k.endswith("running_var")}
            stat_err = max((rs[k] - stats[k]).abs().max().item() for k in rs)
            good = worst >= (0.999999 if precision == "fp32" else 0.99) and stat_err < tol * 10
            ok &= good
            print(f"[dp_check] {precision}: world {world}, buckets all-reduced {sync.launched}, min grad cosine vs 1-rank {worst:.7f}, "
                  f"running_var max err {stat_err:.2e} -> {'OK' if good else 'BAD'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
