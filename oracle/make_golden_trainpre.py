"""Generate tests/golden/trainpre.pt by running the UNMODIFIED reference preprocessing (TrainPre, real cv2) -- TEST INFRASTRUCTURE.
Run in the build container only (needs /root/reference and opencv):   python oracle/make_golden_trainpre.py"""
import contextlib
import io
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shim"), "/root/reference"]
with contextlib.redirect_stdout(io.StringIO()):
    from utils.dataloader.dataloader import TrainPre  # noqa: E402

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]          # local_configs/NYUDepthv2/*.py
SCALES = [0.5, 0.75, 1, 1.25, 1.5, 1.75]


def main():
    cases = []
    rng = np.random.default_rng(0)
    # seeds chosen to cover: mirror on/off, every scale (0.5 = exact-2x INTER_AREA path + padding, 0.75, 1 = copy, up-scaling), crops at
    # the image end, odd source sizes (generic path at scale 0.5), the `sign` normalisation of the depth image, no scale array
    for (H, W, ch, cw, sign, scales, seeds) in [(80, 112, 48, 64, False, SCALES, (0, 1, 3, 6, 7, 4)), (80, 112, 48, 64, True, SCALES, (2, 11)),
                                                (61, 75, 32, 40, False, SCALES, (0, 1, 5, 6)), (40, 48, 40, 48, False, None, (0, 3))]:
        cfg = SimpleNamespace(train_scale_array=scales, image_height=ch, image_width=cw)
        pre = TrainPre(MEAN, STD, sign=sign, config=cfg)
        yy, xx = np.mgrid[0:H, 0:W]
        for seed in seeds:
            noise = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            smooth = np.stack([(yy * 255 // (H - 1)), (xx * 255 // (W - 1)), ((3 * yy + 2 * xx) % 256)], -1).astype(np.uint8)
            rgb = noise if seed % 2 == 0 else smooth
            modal = np.ascontiguousarray(rng.integers(0, 256, (H, W, 1), dtype=np.uint8).repeat(3, axis=2))
            gt = rng.integers(0, 41, (H, W), dtype=np.uint8)
            random.seed(1000 + seed)
            p_rgb, p_gt, p_modal = pre(rgb.copy(), gt.copy(), modal.copy())
            cases.append(dict(H=H, W=W, crop=(ch, cw), sign=sign, scales=scales, seed=1000 + seed, rgb=torch.from_numpy(rgb), gt=torch.from_numpy(gt),
                              modal=torch.from_numpy(modal), out_rgb=torch.from_numpy(np.ascontiguousarray(p_rgb)).float(),
                              out_gt=torch.from_numpy(np.ascontiguousarray(p_gt)), out_modal=torch.from_numpy(np.ascontiguousarray(p_modal)).float()))
    torch.save(dict(mean=MEAN, std=STD, cases=cases), os.path.join(ROOT, "tests", "golden", "trainpre.pt"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
