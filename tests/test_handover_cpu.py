"""Host logic of the cross-Block residual hand-over (dformer_b200/functions.py `_carried_layernorm`, `_mlp_fwd`'s tail) with the
kernel launchers replaced by CPU stand-ins: a pending residual is consumed exactly once, by the Block whose input buffer it
fills, and a foreign buffer is refused."""
from types import SimpleNamespace

import pytest
import torch

from dformer_b200 import functions as Fn


class _FakeK:
    def __init__(self):
        self.calls = []

    def layernorm_fwd(self, x, g, b, eps, dtype):
        self.calls.append("ln")
        y = torch.nn.functional.layer_norm(x, (x.shape[1],), g, b, eps).to(dtype)
        return y, x.mean(1), x.var(1, unbiased=False).add(eps).rsqrt()

    def scale_residual_layernorm_fwd(self, res, branch, ls, scale_b, rows_per_sample, g, b, eps, out=None):
        self.calls.append("res+ln")
        x1 = res + ls * branch.float()
        out.copy_(x1)
        y = torch.nn.functional.layer_norm(x1, (x1.shape[1],), g, b, eps).to(branch.dtype)
        return out, y, x1.mean(1), x1.var(1, unbiased=False).add(eps).rsqrt()


@pytest.fixture
def fake_k(monkeypatch):
    k = _FakeK()
    monkeypatch.setattr(Fn.K, "layernorm_fwd", k.layernorm_fwd)
    monkeypatch.setattr(Fn.K, "scale_residual_layernorm_fwd", k.scale_residual_layernorm_fwd)
    return k


def test_pending_residual_is_formed_in_the_next_blocks_input_buffer_and_consumed_once(fake_k):
    torch.manual_seed(0)
    M, C = 12, 8
    res, f, ls = torch.randn(M, C), torch.randn(M, C), torch.full((C,), 0.5)
    g, b = torch.ones(C), torch.zeros(C)
    out = torch.empty(M, C)                                  # what the previous Block returned (still unwritten)
    pending = {"mlp.": (res, f, ls, None, out)}
    st = SimpleNamespace(dtype=torch.float32, H=3, W=4, pending=pending)
    xn, mu, rs = Fn._carried_layernorm(out, "mlp.", st, g, b)
    assert fake_k.calls == ["res+ln"] and not pending
    torch.testing.assert_close(out, res + 0.5 * f)
    torch.testing.assert_close(xn, torch.nn.functional.layer_norm(res + 0.5 * f, (C,), g, b, 1e-6))
    # nothing pending any more: the plain LayerNorm runs
    xn2, _, _ = Fn._carried_layernorm(out, "mlp.", st, g, b)
    assert fake_k.calls == ["res+ln", "ln"]
    torch.testing.assert_close(xn2, xn)


def test_a_pending_residual_for_another_buffer_is_refused(fake_k):
    M, C = 4, 8
    res, f, ls = torch.randn(M, C), torch.randn(M, C), torch.ones(C)
    st = SimpleNamespace(dtype=torch.float32, H=2, W=2, pending={"mlp_e2.": (res, f, ls, None, torch.empty(M, C))})
    with pytest.raises(AssertionError, match="pending residual"):
        Fn._carried_layernorm(torch.empty(M, C), "mlp_e2.", st, torch.ones(C), torch.zeros(C))


def test_states_without_a_hand_over_take_the_plain_path(fake_k):
    st = SimpleNamespace(dtype=torch.float32, H=2, W=2)     # e.g. a Block driven directly by a kernel test
    Fn._carried_layernorm(torch.randn(4, 8), "mlp.", st, torch.ones(8), torch.zeros(8))
    assert fake_k.calls == ["ln"]
