"""ORACLE (test infrastructure, not product code).

CPU restatement, in float64, of the algorithm of the one-pass fused step (`cosinesampler_b200/csrc/cs_fused.cuh`,
`fused.py`) -- the formulas the CUDA kernels evaluate, written with plain torch ops on top of the jet oracle
(`oracle/stage_oracle.py`), so that the algebra of the kernel is pinned on the CPU against the reference's own
formulation (nested autograd over its pure-PyTorch sampler and the Linear-Tanh-Linear head, `test/test_2d.py:36-127`):

    premix    Vh[n, k, t] = sum_c W1[k, c] V[n, c, t]                      (cs_head_premix)
    gather    H_j[k, p]   = jets of Vh at the points: value, d/dx_a, d2/dx_a^2, summed over the cells
                            (the sampler and the first layer commute)       (cs_pde_fused_kernel, phase A)
    head      t = tanh(H_0 + b1), s1 = 1 - t^2, s2 = -2 t s1, s3 = -2 (s1^2 + t s2)
              u = w2.t + b2, u_a = w2.(s1 H_a), u_aa = w2.(s2 H_a^2 + s1 H_aa)
              f = c_u u + c_u3 u^3 + sum_a c1_a u_a + c2_a u_aa, loss = scale sum_p f^2
              gg = 2 scale f, gsc = gg (c_u + 3 c_u3 u^2), g1_a = gg c1_a, g2_a = gg c2_a
              G1 = sum_a g1_a H_a + g2_a H_aa, G2 = sum_a g2_a H_a^2
              d loss / d H_0  = w2 (gsc s1 + s2 G1 + s3 G2)
              d loss / d H_a  = w2 (s1 g1_a + 2 s2 g2_a H_a)
              d loss / d H_aa = w2 s1 g2_a
              gb1 = sum_p d loss / d H_0, gw2 = sum_p (gsc t + s1 G1 + s2 G2), gb2 = sum_p gsc   (phase B)
    scatter   gVh = adjoint of the gather applied to d loss / d H_j        (phase C)
    postmix   gInput[n, c, t] = sum_k W1[k, c] gVh[n, k, t],  gW1[k, c] = sum_{n,t} gVh[n, k, t] V[n, c, t]
                                                                            (cs_head_postmix)

Pinned by tests/test_onepass_oracle.py.  Only tests/ may import this module.
"""
import torch

from . import stage_oracle as so


def one_pass_step(cells, coords, W1, b1, w2, b2, coef, scale, offset, kernel=so.K_COSINE, pad=so.PAD_ZEROS,
                  align=True, multicell=True, index_mode=2):
    """cells [N,C,*S], coords [P,dim], head parameters W1 [K,C], b1 [K], w2 [K], b2 []; coef = dict(c_u, c_u3,
    c1 [dim], c2 [dim]); loss = scale * sum_p f^2.
    -> dict(loss, gInput [N,C,*S], gW1, gb1, gw2, gb2, u [P], f [P])"""
    dt = torch.float64
    V = cells.to(dt)
    N, C = V.shape[:2]
    dim = coords.shape[-1]
    W1, b1, w2, b2 = W1.to(dt), b1.to(dt), w2.reshape(-1).to(dt), b2.reshape(()).to(dt)
    kw = dict(pad=pad, align=align, kernel=kernel, multicell=multicell, index_mode=index_mode)
    # premix + gather
    Vh = torch.einsum("kc,nc...->nk...", W1, V)
    H = so.jet_forward(Vh, coords.to(dt), offset, order=2, **kw)                 # [1 + 2 dim, K, P]
    H0, Ha, Haa = H[0], H[1:1 + dim], H[1 + dim:1 + 2 * dim]
    # head
    t = torch.tanh(H0 + b1[:, None])
    s1 = 1 - t * t
    s2 = -2 * t * s1
    s3 = -2 * (s1 * s1 + t * s2)
    w = w2[:, None]
    u = (w * t).sum(0) + b2
    ua = (w * s1 * Ha).sum(1)                                                   # [dim, P]
    uaa = (w * (s2 * Ha * Ha + s1 * Haa)).sum(1)
    c1 = torch.tensor([float(x) for x in coef["c1"]][:dim], dtype=dt)
    c2 = torch.tensor([float(x) for x in coef["c2"]][:dim], dtype=dt)
    c_u, c_u3 = float(coef["c_u"]), float(coef["c_u3"])
    f = c_u * u + c_u3 * u ** 3 + (c1[:, None] * ua + c2[:, None] * uaa).sum(0)
    loss = scale * (f * f).sum()
    gg = 2 * scale * f
    gsc = gg * (c_u + 3 * c_u3 * u * u)
    g1 = gg[None, :] * c1[:, None]                                              # [dim, P]
    g2 = gg[None, :] * c2[:, None]
    G1 = (g1[:, None, :] * Ha + g2[:, None, :] * Haa).sum(0)                    # [K, P]
    G2 = (g2[:, None, :] * Ha * Ha).sum(0)
    gH = torch.zeros_like(H)
    gH[0] = w * (gsc * s1 + s2 * G1 + s3 * G2)
    gH[1:1 + dim] = w * (s1 * g1[:, None, :] + 2 * s2 * g2[:, None, :] * Ha)
    gH[1 + dim:] = w * s1 * g2[:, None, :]
    gb1 = gH[0].sum(1)
    gw2 = (gsc * t + s1 * G1 + s2 * G2).sum(1)
    gb2 = gsc.sum()
    # scatter + postmix
    gVh = so.jet_backward(gH, Vh.shape, coords.to(dt), offset, order=2, **kw)
    gInput = torch.einsum("kc,nk...->nc...", W1, gVh)
    gW1 = torch.einsum("nk...,nc...->kc", gVh, V)
    return dict(loss=loss, gInput=gInput, gW1=gW1, gb1=gb1, gw2=gw2, gb2=gb2, u=u, f=f)
