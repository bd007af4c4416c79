"""ncu probe: the BatchNorm streaming kernels at the LightHamHead shape (M = 38400, C = 512, bf16), eager launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dformer_b200 import kernels as K  # noqa: E402

M, C, B = 38400, 512, 8
x = torch.randn(M, C, device="cuda").bfloat16()
dy, res = torch.randn(M, C, device="cuda").bfloat16(), torch.randn(M, C, device="cuda").bfloat16()
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
for _ in range(3):
    st = K.bn_stats(x)
    ms = K.bn_finalize(st, M, 1e-5, 0.1, rm, rv)
    y = K.bn_apply(x, ms, g, b, torch.bfloat16, residual=res, act=K.ACT_RELU)
    gbuf, sums = K.bn_bwd_reduce(dy, x, ms, g, b, res, K.ACT_RELU, None, M // B)
    dx = K.bn_bwd_apply(gbuf, x, ms, g, sums, M, True, torch.bfloat16)
torch.cuda.synchronize()
