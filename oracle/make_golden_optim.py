"""Generate tests/golden/optim.json by running the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only):
    python oracle/make_golden_optim.py
Pins row N1 of SURVEY.md 8(f): the two AdamW parameter groups `group_weight` builds for every DFormer variant
(utils/init_func.py:26-70 as called by utils/train.py:207-209) -- recorded by parameter NAME in the reference's own order, which is also
`torch.optim.AdamW`'s state-dict numbering -- and `WarmUpPolyLR` (utils/lr_policy.py:22-34) sampled over a schedule."""
import contextlib
import io
import json
import os
import sys

import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
from make_golden import OUT, build_reference  # noqa: E402  (puts the mmcv/mmengine stand-ins and /root/reference on sys.path)


def main():
    from utils.init_func import group_weight
    from utils.lr_policy import WarmUpPolyLR
    gold = {"groups": {}, "lr": []}
    for name, ncls in (("DFormer-Tiny", 40), ("DFormer-Small", 40), ("DFormer-Base", 37), ("DFormer-Large", 40)):
        m = build_reference(name, ncls)
        names = {id(p): k for k, p in m.named_parameters()}
        with contextlib.redirect_stdout(io.StringIO()):
            groups = group_weight([], m, nn.BatchNorm2d, 6e-5)
        gold["groups"][name] = {"decay": [names[id(p)] for p in groups[0]["params"]], "no_decay": [names[id(p)] for p in groups[1]["params"]],
                                "no_decay_weight_decay": groups[1]["weight_decay"], "n_parameters": len(names)}
    # NYUDepthv2 DFormer-L schedule (local_configs: lr 6e-5, power 0.9, 500 epochs x niters, 10 warm-up epochs) + small edge cases
    for start, power, total, warm in ((6e-5, 0.9, 500 * 99, 10 * 99), (8e-5, 1.0, 1000, 0), (1e-3, 0.9, 50, 50)):
        pol = WarmUpPolyLR(start, power, total, warm)
        its = sorted({0, 1, 2, warm - 1, warm, warm + 1, total // 2, total - 1} & set(range(total)))
        gold["lr"].append({"start_lr": start, "lr_power": power, "total_iters": total, "warmup_steps": warm,
                           "samples": [[i, pol.get_lr(i)] for i in its]})
    with open(os.path.join(OUT, "optim.json"), "w") as f:
        json.dump(gold, f)
    print({k: (len(v["decay"]), len(v["no_decay"]), v["n_parameters"]) for k, v in gold["groups"].items()})


if __name__ == "__main__":
    main()
