"""ORACLE (test infrastructure, not product code).

CPU restatement of what each of the reference's four CUDA entry points
computes (`forward`, `backward`, `backward_backward`,
`backward_backward_backward`; pybind table `cosine_sampler_2d.cpp:130-135`,
`cosine_sampler_3d.cpp:133-138`), for 2D and 3D, written as vectorised torch
ops on the CPU.  It is *bug-compatible* with the CUDA kernels:

  * 2D forward always maps coordinates with align_corners=True
    (`cosine_sampler_2d_kernel.cu:307-308`); 3D honours the flag (`cu3d:299-301`);
  * 2D double backward has no mixed second derivatives and no
    gOutInput -> gGrid term (`cu2d:675-678,705-706`); 3D has both (`cu3d:836-856`);
  * triple backward keeps only the pure second-derivative terms, 2D and 3D
    (`cu2d:876-885`, `cu3d:1054-1065`), and ignores gOutgInput;
  * reflection padding with align_corners reflects over [0, S-2]
    (`cu2d:184-188,226-230`), not ATen's [0, S-1].

Index map and fractional position are evaluated in fp32 exactly as the fp32
kernels do (so cell decisions are identical); everything after that runs in
`compute_dtype` (fp64 by default), which makes this a tight reference for fp32
kernels.  index_mode 0 = multiply and add rounded separately (what the
pure-PyTorch sampler does, `test/grid_sampler.py:37-38`); index_mode 1 = one
fused multiply-add (what the reference's `--use_fast_math` build emits, SURVEY
section 7.1); index_mode 2 = no fp32 rounding at all (used to pin the formulas
against fp64 autograd).

Pinned by `tests/test_stage_oracle.py`: F, B, BB, BBB here agree with PyTorch
autograd over `oracle/grid_sampler_oracle.py` (itself bit-equal to the real
reference on the golden vectors) wherever the CUDA kernels are exact, and on
the GPU box against the recompiled reference CUDA op when `oracle/_ref` exists.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.
"""
import math

import torch

PAD_ZEROS, PAD_BORDER, PAD_REFLECTION = 0, 1, 2
K_COSINE, K_LINEAR, K_SMOOTHSTEP = 0, 1, 2


def _reflect(i, twice_low, twice_high):
    """reflect_coordinates_set_grad, cu2d:144-171."""
    if twice_low == twice_high:
        return torch.zeros_like(i), torch.zeros_like(i)
    lo = twice_low / 2.0
    span = (twice_high - twice_low) / 2.0
    x = i - lo
    sign = torch.where(x < 0, -torch.ones_like(x), torch.ones_like(x))
    x = x.abs()
    extra = torch.fmod(x, span)
    flips = torch.floor(x / span).to(torch.int64)
    even = (flips % 2) == 0
    out = torch.where(even, extra + lo, span - extra + lo)
    grad = torch.where(even, sign, -sign)
    return out, grad


def _clip(i, size):
    """clip_coordinates_set_grad, cu2d:98-116."""
    hi = float(size - 1)
    g = torch.where((i <= 0) | (i >= hi), torch.zeros_like(i), torch.ones_like(i))
    return torch.clamp(i, 0.0, hi), g


def axis_terms(g, size, off, pad, align, kernel, multicell, index_mode, compute_dtype):
    """Per-axis quantities of SURVEY section 7.0 for coordinates g [N,P] (fp32)
    and offsets off [N,1] (fp32).  Returns l (int64), W, dW, d2W (each a pair
    low/high, compute_dtype) and inb (pair of bool masks)."""
    if index_mode == 2:          # exact-arithmetic pin: no fp32 rounding anywhere
        g = g.to(compute_dtype)
        off = off.to(compute_dtype)
    else:
        g = g.to(torch.float32)
        off = off.to(torch.float32)
    if align:
        s = size - 1 - (1 if multicell else 0)      # cu2d:56-61
        m = s / 2.0                                  # cu2d:80
        if index_mode in (0, 2):
            i = ((g + 1) / 2) * float(s) + off
        else:
            i = (((g + 1) / 2).double() * float(s) + off.double()).float()
    else:
        m = size / 2.0                               # cu2d:84
        if index_mode in (0, 2):
            i = ((g + 1) * float(size) - 1) / 2 + off
        else:
            i = (((g + 1).double() * float(size) - 1) / 2 + off.double()).float()
    mult = torch.full_like(i, m)
    if pad == PAD_BORDER:
        i, gc = _clip(i, size)
        mult = mult * gc
    elif pad == PAD_REFLECTION:
        if align:
            i, gr = _reflect(i, 0, 2 * (size - 2))
        else:
            i, gr = _reflect(i, -1, 2 * size - 1)
        i, gc = _clip(i, size)
        mult = mult * gr * gc
    lf = torch.floor(i)
    l = lf.to(torch.int64)
    r = ((lf + 1) - i).to(compute_dtype)             # fp32 subtraction, as the kernels
    t = (i - lf).to(compute_dtype)
    mult = mult.to(compute_dtype)
    if kernel == K_COSINE:
        k0 = 0.5 * (1 - torch.cos(math.pi * r))
        k1 = 0.5 * math.pi * torch.sin(math.pi * r)
        k2 = 0.5 * math.pi ** 2 * torch.cos(math.pi * r)
        w_hi = 1 - k0
    elif kernel == K_SMOOTHSTEP:
        k0 = r * r * (3 - 2 * r)
        k1 = 6 * r * (1 - r)
        k2 = 6 - 12 * r
        w_hi = 1 - k0
    else:
        k0, k1, k2 = r, torch.ones_like(r), torch.zeros_like(r)
        w_hi = t
    W = (k0, w_hi)
    dW = (-mult * k1, mult * k1)
    d2W = (mult * mult * k2, -mult * mult * k2)
    inb = ((l >= 0) & (l < size), (l + 1 >= 0) & (l + 1 < size))
    return l, W, dW, d2W, inb


class _Setup:
    def __init__(self, input, grid, offset, pad, align, kernel, multicell, index_mode,
                 compute_dtype):
        self.nd = nd = grid.shape[-1]
        self.N, self.C = input.shape[:2]
        self.sizes = [input.shape[-1 - a] for a in range(nd)]       # axis 0 -> W, 1 -> H, 2 -> D
        self.P = grid[0].numel() // nd
        g = grid.reshape(self.N, self.P, nd)
        off = offset.reshape(self.N, 1)
        self.ax = [axis_terms(g[..., a], self.sizes[a], off, pad, align, kernel, multicell,
                              index_mode, compute_dtype) for a in range(nd)]
        self.dt = compute_dtype
        strides = [1] * nd
        for a in range(1, nd):
            strides[a] = strides[a - 1] * self.sizes[a - 1]
        self.strides = strides
        self.T = strides[-1] * self.sizes[-1]

    def corners(self):
        """yield (bits, flat_index [N,P] clamped, inb [N,P])"""
        for q in range(1 << self.nd):
            bits = [(q >> a) & 1 for a in range(self.nd)]
            idx = 0
            inb = None
            for a in range(self.nd):
                l = self.ax[a][0] + bits[a]
                idx = idx + torch.clamp(l, 0, self.sizes[a] - 1) * self.strides[a]
                m = self.ax[a][4][bits[a]]
                inb = m if inb is None else (inb & m)
            yield bits, idx, inb

    def w(self, bits):
        out = None
        for a in range(self.nd):
            x = self.ax[a][1][bits[a]]
            out = x if out is None else out * x
        return out

    def d1(self, bits, a):
        out = self.ax[a][2][bits[a]]
        for b in range(self.nd):
            if b != a:
                out = out * self.ax[b][1][bits[b]]
        return out

    def d2(self, bits, a, b):
        if a == b:
            out = self.ax[a][3][bits[a]]
            for c in range(self.nd):
                if c != a:
                    out = out * self.ax[c][1][bits[c]]
            return out
        out = self.ax[a][2][bits[a]] * self.ax[b][2][bits[b]]
        for c in range(self.nd):
            if c != a and c != b:
                out = out * self.ax[c][1][bits[c]]
        return out

    def gather(self, field, idx, inb):
        flat = field.reshape(self.N, self.C, self.T).to(self.dt)
        v = torch.gather(flat, 2, idx.view(self.N, 1, self.P).expand(self.N, self.C, self.P))
        return v * inb.view(self.N, 1, self.P).to(self.dt)

    def scatter(self, acc, idx, inb, vals):
        vals = vals * inb.view(self.N, 1, self.P).to(self.dt)
        acc.scatter_add_(2, idx.view(self.N, 1, self.P).expand(self.N, self.C, self.P), vals)


def _stream(x, N, C, P, dt):
    return x.reshape(N, C, P).to(dt)


def forward(input, grid, offset, pad=0, align=True, kernel=0, multicell=True, index_mode=0,
            compute_dtype=torch.float64):
    """out[n,c,p] = sum_q V_q w_q  (cu2d:265-356, cu3d:250-371)."""
    nd = grid.shape[-1]
    if nd == 2:
        align = True                                   # cu2d:307-308
    s = _Setup(input, grid, offset, pad, align, kernel, multicell, index_mode, compute_dtype)
    out = torch.zeros(s.N, s.C, s.P, dtype=compute_dtype)
    for bits, idx, inb in s.corners():
        out += s.gather(input, idx, inb) * s.w(bits).view(s.N, 1, s.P)
    return out.reshape(input.shape[:2] + grid.shape[1:-1])


def backward(gOut, input, grid, offset, pad=0, align=True, input_requires_grad=True, kernel=0,
             multicell=True, index_mode=0, compute_dtype=torch.float64):
    """gInput[l+q] += w_q gOut;  gGrid_a = sum_c gOut sum_q V_q D_a,q
    (cu2d:359-507, cu3d:373-584)."""
    s = _Setup(input, grid, offset, pad, align, kernel, multicell, index_mode, compute_dtype)
    go = _stream(gOut, s.N, s.C, s.P, s.dt)
    gI = torch.zeros(s.N, s.C, s.T, dtype=s.dt) if input_requires_grad else None
    gG = torch.zeros(s.N, s.P, s.nd, dtype=s.dt)
    for bits, idx, inb in s.corners():
        if gI is not None:
            s.scatter(gI, idx, inb, go * s.w(bits).view(s.N, 1, s.P))
        S = (s.gather(input, idx, inb) * go).sum(1)
        for a in range(s.nd):
            gG[..., a] += S * s.d1(bits, a)
    if gI is not None:
        gI = gI.reshape(input.shape)
    return gI, gG.reshape(grid.shape)


def backward_backward(gOutInput, gOutGrid, input, grid, gOut, offset, pad=0, align=True,
                      input_requires_grad=False, kernel=0, multicell=True, index_mode=0,
                      compute_dtype=torch.float64):
    """(gInput, gGrid, ggOut) of cu2d:509-717 / cu3d:587-870.
    gOutInput is only read when input_requires_grad (the flag of mod2d:87)."""
    s = _Setup(input, grid, offset, pad, align, kernel, multicell, index_mode, compute_dtype)
    nd = s.nd
    go = _stream(gOut, s.N, s.C, s.P, s.dt)
    gog = gOutGrid.reshape(s.N, s.P, nd).to(s.dt)
    use_goi = bool(input_requires_grad) and gOutInput is not None
    gI = torch.zeros(s.N, s.C, s.T, dtype=s.dt)
    gG = torch.zeros(s.N, s.P, nd, dtype=s.dt)
    ggO = torch.zeros(s.N, s.C, s.P, dtype=s.dt)
    for bits, idx, inb in s.corners():
        V = s.gather(input, idx, inb)
        A = sum(s.d1(bits, a) * gog[..., a] for a in range(nd))           # [N,P]
        ggO += V * A.view(s.N, 1, s.P)
        s.scatter(gI, idx, inb, go * A.view(s.N, 1, s.P))
        S = (V * go).sum(1)
        if use_goi:
            U = s.gather(gOutInput, idx, inb)
            ggO += U * s.w(bits).view(s.N, 1, s.P)
            Tq = (U * go).sum(1)
        for a in range(nd):
            if nd == 2:
                gG[..., a] += S * s.d2(bits, a, a) * gog[..., a]
            else:
                gG[..., a] += S * sum(s.d2(bits, a, b) * gog[..., b] for b in range(nd))
                if use_goi:
                    gG[..., a] += Tq * s.d1(bits, a)
    return gI.reshape(input.shape), gG.reshape(grid.shape), ggO.reshape(gOut.shape)


def backward_backward_backward(input, grid, gOut, gOutGrid, gOutgGrid, offset, pad=0, align=True,
                               input_requires_grad=False, kernel=0, multicell=True, index_mode=0,
                               compute_dtype=torch.float64):
    """(gInput, ggOut) of cu2d:722-891 / cu3d:875-1071 (pure second derivatives)."""
    s = _Setup(input, grid, offset, pad, align, kernel, multicell, index_mode, compute_dtype)
    nd = s.nd
    go = _stream(gOut, s.N, s.C, s.P, s.dt)
    gog = gOutGrid.reshape(s.N, s.P, nd).to(s.dt)
    gogg = gOutgGrid.reshape(s.N, s.P, nd).to(s.dt)
    gI = torch.zeros(s.N, s.C, s.T, dtype=s.dt)
    ggO = torch.zeros(s.N, s.C, s.P, dtype=s.dt)
    for bits, idx, inb in s.corners():
        V = s.gather(input, idx, inb)
        E = sum(s.d2(bits, a, a) * gogg[..., a] * gog[..., a] for a in range(nd))
        ggO += V * E.view(s.N, 1, s.P)
        s.scatter(gI, idx, inb, go * E.view(s.N, 1, s.P))
    return gI.reshape(input.shape), ggO.reshape(gOut.shape)


# ---------------------------------------------------------------------------
# Jet operator (cosinesampler_b200/jet.py; not in the reference): the same per-corner terms as
# the four entry points above, summed over the cells and kept per channel.  Pinned in
# tests/test_stage_oracle.py against forward / backward / backward_backward /
# backward_backward_backward above (contracting the channels with a random gOut).
# ---------------------------------------------------------------------------
def _jet_setup(input, coords, offset, pad, align, kernel, multicell, index_mode, compute_dtype):
    N = input.shape[0]
    nd = coords.shape[-1]
    P = coords.shape[0]
    grid = coords.reshape((1,) * nd + (P, nd)).expand((N,) + (1,) * (nd - 1) + (P, nd))
    return _Setup(input, grid, offset, pad, align, kernel, multicell, index_mode, compute_dtype)


def jet_count(nd, order):
    """Number of jets: value + nd first + (order >= 2) nd pure second + (order == 3) nd(nd-1)/2 mixed second
    derivatives, in the order (x,y) in 2D and (x,y), (x,z), (y,z) in 3D."""
    return 1 + min(order, 2) * nd + (nd * (nd - 1) // 2 if order >= 3 else 0)


def mixed_pairs(nd):
    return [(a, b) for a in range(nd) for b in range(a + 1, nd)]


def jet_forward(input, coords, offset, order=2, pad=0, align=True, kernel=0, multicell=True,
                index_mode=0, compute_dtype=torch.float64):
    """jets [jet_count(dim, order), C, P]: value, d/dg_a, d2/dg_a^2 (and, order 3, d2/dg_a dg_b) of
    sum_n sample(input[n], coords).  The mixed terms are what the reference's 3D double backward contracts
    (cu3d:836-856); its 2D kernels leave them out (cu2d:675-678)."""
    s = _jet_setup(input, coords, offset, pad, align, kernel, multicell, index_mode, compute_dtype)
    nd = s.nd
    jets = torch.zeros(jet_count(nd, order), s.C, s.P, dtype=s.dt)
    for bits, idx, inb in s.corners():
        V = s.gather(input, idx, inb)                                   # [N,C,P]
        jets[0] += (V * s.w(bits).view(s.N, 1, s.P)).sum(0)
        for a in range(nd):
            jets[1 + a] += (V * s.d1(bits, a).view(s.N, 1, s.P)).sum(0)
            if order >= 2:
                jets[1 + nd + a] += (V * s.d2(bits, a, a).view(s.N, 1, s.P)).sum(0)
        if order >= 3:
            for m, (a, b) in enumerate(mixed_pairs(nd)):
                jets[1 + 2 * nd + m] += (V * s.d2(bits, a, b).view(s.N, 1, s.P)).sum(0)
    return jets


def jet_backward(gJets, input_shape, coords, offset, order=2, pad=0, align=True, kernel=0,
                 multicell=True, index_mode=0, compute_dtype=torch.float64):
    """Adjoint of jet_forward: gInput [N,C,*S]."""
    proto = torch.zeros(input_shape, dtype=compute_dtype)
    s = _jet_setup(proto, coords, offset, pad, align, kernel, multicell, index_mode, compute_dtype)
    nd = s.nd
    g = gJets.to(s.dt)
    gI = torch.zeros(s.N, s.C, s.T, dtype=s.dt)
    for bits, idx, inb in s.corners():
        coef = g[0].unsqueeze(0) * s.w(bits).view(s.N, 1, s.P)
        for a in range(nd):
            coef = coef + g[1 + a].unsqueeze(0) * s.d1(bits, a).view(s.N, 1, s.P)
            if order >= 2:
                coef = coef + g[1 + nd + a].unsqueeze(0) * s.d2(bits, a, a).view(s.N, 1, s.P)
        if order >= 3:
            for m, (a, b) in enumerate(mixed_pairs(nd)):
                coef = coef + g[1 + 2 * nd + m].unsqueeze(0) * s.d2(bits, a, b).view(s.N, 1, s.P)
        s.scatter(gI, idx, inb, coef)
    return gI.reshape(input_shape)
