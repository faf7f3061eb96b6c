"""ORACLE (test infrastructure, not product code).

The caller's head of the reference's scripts, Linear(C,K)-Tanh-Linear(K,1) (`test/test_2d.py:42-47`), as a
chain of three `torch.autograd.Function`s with closed-form backward passes -- what a fused, triple-
differentiable head would have to compute at each level of the reference's nested-autograd chain
(`test_2d.py:55-127`):

    level 0   u      = w2 . tanh(W1 z + b1) + b2
    level 1   gz     = d <gu, u> / dz                            (+ parameter gradients)
    level 2   (gz2, ggu) = d <ggz, gz> / d(z, gu)                (+ parameter gradients)
    level 3   d (<g_gz2, gz2> + <g_ggu, ggu>) / d(z, gu, ggz, parameters)

DESIGN.md section 9 names this as the next step for the drop-in arm (the head under torch autograd is
5.5 of the 8.2 ms per 2^20 points).  Nothing in the product imports this module; it is pinned to PyTorch
autograd over the plain nn.Sequential head by tests/test_head_oracle.py and serves as the specification
(and future checker) of those kernels.

Notation per point: h = W1 z + b1, t = tanh h, s1 = 1 - t^2, s2 = -2 t s1, s3 = -2 (s1^2 + t s2).
"""
import torch


def _act(z, W1, b1):
    h = z @ W1.t() + b1
    t = torch.tanh(h)
    s1 = 1 - t * t
    s2 = -2 * t * s1
    s3 = -2 * (s1 * s1 + t * s2)
    return t, s1, s2, s3


def level0(z, W1, b1, w2, b2):
    """u [P]"""
    t, _, _, _ = _act(z, W1, b1)
    return t @ w2.reshape(-1) + b2.reshape(())


def level1(z, gu, W1, b1, w2):
    """gz [P,C] = d<gu,u>/dz and the parameter gradients (gW1 [K,C], gb1 [K], gw2 [K]) of <gu,u>."""
    t, s1, _, _ = _act(z, W1, b1)
    a = gu[:, None] * w2.reshape(1, -1) * s1                       # [P,K]
    return a @ W1, a.t() @ z, a.sum(0), (gu[:, None] * t).sum(0)


def level2(z, gu, ggz, W1, b1, w2):
    """S = <ggz, gz>:  gz2 = dS/dz [P,C], ggu = dS/dgu [P], and dS/d(W1, b1, w2)."""
    t, s1, s2, _ = _act(z, W1, b1)
    w = w2.reshape(1, -1)
    v = ggz @ W1.t()                                               # [P,K]
    q = gu[:, None] * w * s2 * v
    a = gu[:, None] * w * s1
    return (q @ W1, (w * s1 * v).sum(1), q.t() @ z + a.t() @ ggz, q.sum(0), (gu[:, None] * s1 * v).sum(0))


def level3(z, gu, ggz, g_gz2, g_ggu, W1, b1, w2):
    """T = <g_gz2, gz2> + <g_ggu, ggu>:  dT/d(z, gu, ggz, W1, b1, w2)."""
    t, s1, s2, s3 = _act(z, W1, b1)
    w = w2.reshape(1, -1)
    v = ggz @ W1.t()
    r = g_gz2 @ W1.t()
    m = gu[:, None] * s2 * r + g_ggu[:, None] * s1
    n = gu[:, None] * s3 * r + g_ggu[:, None] * s2
    wvn = w * v * n
    wm = w * m
    g_z = wvn @ W1
    g_gu = (w * v * s2 * r).sum(1)
    g_ggz = wm @ W1
    g_W1 = wvn.t() @ z + wm.t() @ ggz + (w * v * gu[:, None] * s2).t() @ g_gz2
    return g_z, g_gu, g_ggz, g_W1, wvn.sum(0), (v * m).sum(0)


class _L2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, gu, ggz, W1, b1, w2):
        ctx.save_for_backward(z, gu, ggz, W1, b1, w2)
        ctx.set_materialize_grads(False)
        return level2(z, gu, ggz, W1, b1, w2)

    @staticmethod
    def backward(ctx, g_gz2, g_ggu, gW, gb, gw):
        if gW is not None or gb is not None or gw is not None:
            raise NotImplementedError("gradients of the level-2 parameter gradients")
        z, gu, ggz, W1, b1, w2 = ctx.saved_tensors
        if g_gz2 is None and g_ggu is None:
            return None, None, None, None, None, None
        g_gz2 = torch.zeros_like(z) if g_gz2 is None else g_gz2
        g_ggu = torch.zeros_like(gu) if g_ggu is None else g_ggu
        g_z, g_gu, g_ggz, g_W1, g_b1, g_w2 = level3(z, gu, ggz, g_gz2, g_ggu, W1, b1, w2)
        return g_z, g_gu, g_ggz, g_W1, g_b1, g_w2.reshape(w2.shape)


class _L1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, gu, W1, b1, w2):
        ctx.save_for_backward(z, gu, W1, b1, w2)
        ctx.set_materialize_grads(False)
        gz, gW1, gb1, gw2 = level1(z, gu, W1, b1, w2)
        return gz, gW1, gb1, gw2.reshape(w2.shape)

    @staticmethod
    def backward(ctx, ggz, gW, gb, gw):
        if gW is not None or gb is not None or gw is not None:
            raise NotImplementedError("gradients of the level-1 parameter gradients")
        z, gu, W1, b1, w2 = ctx.saved_tensors
        if ggz is None:
            return None, None, None, None, None
        gz2, ggu, gW1, gb1, gw2 = _L2.apply(z, gu, ggz, W1, b1, w2)
        return gz2, ggu, gW1, gb1, gw2.reshape(w2.shape)


class _L0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, W1, b1, w2, b2):
        ctx.save_for_backward(z, W1, b1, w2)
        return level0(z, W1, b1, w2, b2).reshape(-1, 1)

    @staticmethod
    def backward(ctx, gu):
        z, W1, b1, w2 = ctx.saved_tensors
        gu = gu.reshape(-1)
        gz, gW1, gb1, gw2 = _L1.apply(z, gu, W1, b1, w2)
        return gz, gW1, gb1, gw2, gu.sum().reshape(1)


class ClosedFormHead(torch.nn.Module):
    """Drop-in for `nn.Sequential(Linear(C,K), Tanh(), Linear(K,1))` sharing its parameters."""

    def __init__(self, sequential):
        super().__init__()
        self.seq = sequential

    def forward(self, z):
        l0, l2 = self.seq[0], self.seq[2]
        return _L0.apply(z.contiguous(), l0.weight, l0.bias, l2.weight, l2.bias)
