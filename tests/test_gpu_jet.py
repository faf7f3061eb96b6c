"""The opt-in fused jet operator (`cosinesampler_b200.jet`, SURVEY 8f ranks 1 + 2): one gather pass
for value + first + pure second derivatives summed over the cells, one scatter pass back.

  * kernels vs the jet oracle (itself pinned to the four stage oracles on the CPU);
  * the jets equal what the drop-in operator's forward / backward / double backward return;
  * a PDE step through the jets gives the same u, u_a, u_aa, loss and d loss / d cells as the
    reference's nested-autograd chain through the oracle sampler (fp32 on the GPU, fp64 on the CPU).
"""
import pytest
import torch

from oracle import stage_oracle as so
from oracle.grid_sampler_oracle import derivative_chain, grid_sample_2d, grid_sample_3d, make_head
from util import assert_close_scaled, safe_coords

pytestmark = pytest.mark.gpu

KCODE = {"cosine": 0, "linear": 1, "smooth-step": 2}


def _off(N, multicell):
    from cosinesampler_b200.autograd import cell_offsets
    return cell_offsets(N, multicell, torch.device("cpu")).clone()


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("C", [4, 8, 16, 32])
@pytest.mark.parametrize("dim", [2, 3])
def test_jet_kernels_match_oracle(cuda, dim, C, order):
    from cosinesampler_b200 import jet
    gen = torch.Generator().manual_seed(10 * C + dim + order)
    N = 5
    P = 1003 if C != 16 else 1000                 # ragged and vector-aligned row lengths
    sizes = (9, 12) if dim == 2 else (6, 7, 8)
    inp = torch.rand((N, C) + sizes, generator=gen)
    coords = torch.rand(P, dim, generator=gen) * 2.6 - 1.3          # some points out of range
    off = _off(N, True)
    J = jet.jet_count(dim, order)
    G = torch.randn(J, C, P, generator=gen)
    for kernel, pad, align in (("cosine", 0, True), ("smooth-step", 1, True), ("linear", 2, True),
                               ("cosine", 0, False)):
        if dim == 2 and not align:
            continue        # jets honour align_corners in 2D; the oracle's 2D forward ignores it (cu2d:307)
        kw = dict(pad=pad, align=align, kernel=KCODE[kernel], multicell=True)
        what = "jet %dD C=%d order=%d %s pad=%d align=%s" % (dim, C, order, kernel, pad, align)
        jets = jet.jet_forward(inp.to(cuda), coords.to(cuda), off.to(cuda), pad, align, KCODE[kernel], True, order)
        assert jets.shape == (J, C, P)
        ref = so.jet_forward(inp, coords, off, order=order, **kw)
        # cell flips between fp32 evaluation orders are impossible here (same fp32 index map)
        for jt in range(J):
            assert_close_scaled(jets[jt], ref[jt], what + " jets[%d]" % jt)
        gI = jet.jet_backward(G.to(cuda), inp.to(cuda), coords.to(cuda), off.to(cuda), pad, align,
                              KCODE[kernel], True, order)
        assert gI.shape == inp.shape
        assert_close_scaled(gI, so.jet_backward(G, inp.shape, coords, off, order=order, **kw), what + " gInput")


@pytest.mark.parametrize("dim", [2, 3])
def test_jets_equal_the_dropin_operator(cuda, dim):
    """z = F.sum(0);  <z_a, gOut> = sum_n gGrid_a of B;  <z_aa, gOut> = sum_n gGrid_a of BB."""
    from cosinesampler_b200 import jet, ops
    gen = torch.Generator().manual_seed(dim)
    N, C, P = 4, 16, 4096
    sizes = (32, 32) if dim == 2 else (12, 12, 12)
    inp = torch.rand((N, C) + sizes, generator=gen).to(cuda)
    coords = (torch.rand(P, dim, generator=gen) * 2 - 1).to(cuda)
    off = _off(N, True).to(cuda)
    grid = coords.reshape((1,) * dim + (P, dim)).repeat((N,) + (1,) * (dim + 1)).contiguous()
    gOut = torch.randn(C, P, generator=gen).to(cuda)
    gOutN = gOut.reshape((1, C) + (1,) * (dim - 1) + (P,)).expand((N, C) + (1,) * (dim - 1) + (P,))
    jets = jet.jet_forward(inp, coords, off, 0, True, 0, True, 2)
    assert_close_scaled(jets[0], ops.forward(inp, grid, off, 0, True, 0, True).reshape(N, C, P).sum(0), "z")
    _, gG = ops.backward(gOutN, inp, grid, off, 0, True, False, 0, True)
    for a in range(dim):
        assert_close_scaled((jets[1 + a] * gOut).sum(0), gG.reshape(N, P, dim)[..., a].sum(0), "z_%d" % a,
                            atol_scale=2e-5)
        hot = torch.zeros_like(grid)
        hot[..., a] = 1.0
        _, gG2, ggO = ops.backward_backward(None, hot, inp, grid, gOutN, off, 0, True, False, 0, True,
                                            want=(False, True, True))
        assert_close_scaled((jets[1 + dim + a] * gOut).sum(0), gG2.reshape(N, P, dim)[..., a].sum(0),
                            "z_%d%d" % (a, a), atol_scale=2e-5)
        assert_close_scaled(jets[1 + a], ggO.reshape(N, C, P).sum(0), "z_%d via ggOut" % a)


def _jet_quantities(S, cells0, coords0, head, residual, device, kernel, multicell):
    """u, u_a, u_aa, loss, dloss through the jet operator."""
    from cosinesampler_b200.chain import _residual
    from cosinesampler_b200.jet import jet_mlp
    dim = coords0.shape[1]
    cells = cells0.to(device).clone().requires_grad_(True)
    jets = S.apply(cells, coords0.to(device), "zeros", True, kernel, multicell)
    u, first, second = jet_mlp(head, jets, dim)
    names = "xyz"[:dim]
    q = {"val": jets[0], "u": u}
    for a in range(dim):
        q["u_" + names[a]] = first[a]
        q["u_%s%s" % (names[a], names[a])] = second[a]
    loss = torch.mean(_residual(u, first, second, residual, torch.pi ** 2) ** 2)
    q["loss"] = loss
    q["dloss"] = torch.autograd.grad(loss, cells)[0]
    return q


@pytest.mark.parametrize("dim,kernel,multicell,residual", [
    (2, "cosine", True, "helmholtz"), (2, "cosine", True, "t2d"), (2, "smooth-step", False, "helmholtz"),
    (2, "bilinear", True, "t2d"), (3, "smooth-step", True, "laplace"), (3, "cosine", True, "laplace"),
    (3, "trilinear", False, "laplace")])
def test_jet_step_matches_the_reference_chain(cuda, dim, kernel, multicell, residual):
    from cosinesampler_b200 import jet
    gen = torch.Generator().manual_seed(200 + dim)
    N, C, P = 4, 16, 3000
    shape = (N, C, 32, 32) if dim == 2 else (N, C, 12, 12, 12)
    sizes = [shape[-1 - a] for a in range(dim)]
    cells0 = torch.rand(shape, generator=gen)
    coords0 = safe_coords(P, dim, sizes, N, multicell, gen, margin=0.01).float()
    head32 = make_head(C, seed=1).to(cuda)
    head64 = make_head(C, seed=1, dtype=torch.float64)
    S = jet.SamplerJet2d if dim == 2 else jet.SamplerJet3d
    ours = _jet_quantities(S, cells0, coords0, head32, residual, cuda, kernel, multicell)

    fn = grid_sample_2d if dim == 2 else grid_sample_3d
    step = {"smooth-step": "smoothstep"}.get(kernel, kernel)
    oracle = lambda c, g: fn(c, g, step=step, offset=multicell)

    def chain(device, dtype, head):
        cells = cells0.to(device=device, dtype=dtype).clone().requires_grad_(True)
        coords = [coords0[:, a:a + 1].to(device=device, dtype=dtype).clone().requires_grad_(True)
                  for a in range(dim)]
        return derivative_chain(oracle, cells, coords, head, residual=residual)

    ref32 = chain(cuda, torch.float32, head32)
    ref64 = chain("cpu", torch.float64, head64)
    for name, a in ours.items():
        r64 = ref64[name].sum(0) if name == "val" else ref64[name]
        r32 = ref32[name].sum(0) if name == "val" else ref32[name]
        a = a.reshape(r64.shape)
        assert_close_scaled(a, r64, "jet %s vs fp64 oracle chain" % name, rtol=1e-4, atol_scale=2e-5)
        assert_close_scaled(a, r32.reshape(a.shape), "jet %s vs fp32 oracle chain" % name, rtol=1e-4,
                            atol_scale=2e-5)
    # and at least as close to the fp64 truth as the fp32 nested-autograd chain is
    for name in ("u_xx", "dloss"):
        t = ref64[name].reshape(-1)
        e_ours = (ours[name].double().cpu().reshape(-1) - t).abs().max()
        e_ref = (ref32[name].double().cpu().reshape(-1) - t).abs().max()
        assert e_ours <= 4 * e_ref + 1e-7 * t.abs().max(), (name, float(e_ours), float(e_ref))


def test_jet_training_step_equals_dropin_training_step(cuda):
    """chain.training_step(jet=True) and the drop-in chain give the same loss and gradients
    (cells and head), chunked or not."""
    from cosinesampler_b200 import jet
    from cosinesampler_b200.chain import training_step
    from cosine_sampler_2d import CosineSampler2d
    gen = torch.Generator().manual_seed(5)
    N, C, P = 4, 16, 8192
    cells0 = torch.rand(N, C, 64, 64, generator=gen)
    coords = [(torch.rand(P, 1, generator=gen) * 2 - 1).to(cuda) for _ in range(2)]
    res = {}
    for mode in ("dropin", "jet", "jet_chunked"):
        cells = torch.nn.Parameter(cells0.clone().to(cuda))
        head = make_head(C, seed=3).to(cuda)
        if mode == "dropin":
            sampler = lambda c, g: CosineSampler2d.apply(c, g, "zeros", True, "cosine", True)
            loss = training_step(sampler, cells, coords, head, "helmholtz")
        else:
            sampler = lambda c, x: jet.SamplerJet2d.apply(c, x, "zeros", True, "cosine", True)
            loss = training_step(sampler, cells, coords, head, "helmholtz", jet=True,
                                 chunk=3000 if mode == "jet_chunked" else None)
        res[mode] = (loss, cells.grad, [p.grad for p in head.parameters()])
    for mode in ("jet", "jet_chunked"):
        assert_close_scaled(res[mode][0], res["dropin"][0], mode + " loss", rtol=1e-4)
        assert_close_scaled(res[mode][1], res["dropin"][1], mode + " cells.grad", rtol=1e-4, atol_scale=2e-5)
        for a, b in zip(res[mode][2], res["dropin"][2]):
            assert_close_scaled(a, b, mode + " head grad", rtol=1e-4, atol_scale=2e-5)


def test_jet_api_errors(cuda):
    from cosinesampler_b200 import jet
    cells = torch.rand(2, 6, 8, 8, device=cuda)
    xy = torch.rand(16, 2, device=cuda)
    with pytest.raises(RuntimeError, match="C in"):
        jet.SamplerJet2d.apply(cells, xy)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        jet.SamplerJet2d.apply(torch.rand(2, 8, 8, 8), xy)
    with pytest.raises(RuntimeError, match="coords must be"):
        jet.SamplerJet2d.apply(torch.rand(2, 8, 8, 8, device=cuda), torch.rand(16, 3, device=cuda))
    with pytest.raises(TypeError):
        jet.SamplerJet2d.apply(torch.rand(2, 8, 8, 8, device=cuda), xy, "zeros", True, "cubic", True)
    # empty point set
    out = jet.SamplerJet2d.apply(torch.rand(2, 8, 8, 8, device=cuda), torch.rand(0, 2, device=cuda))
    assert out.shape == (5, 8, 0)


# ---------------------------------------------------------------------------
# fused head + residual kernel (cs_pde_head_step) and the fused training step
# ---------------------------------------------------------------------------
def _torch_head_reference(jets64, head64, dim, residual, k2, scale):
    """loss, f and every gradient by torch autograd over jet_mlp in fp64 on the CPU."""
    from cosinesampler_b200.chain import _residual
    from cosinesampler_b200.jet import jet_mlp
    jets64 = jets64.clone().requires_grad_(True)
    u, first, second = jet_mlp(head64, jets64, dim)
    f = _residual(u, first, second, residual, k2)
    loss_sum = (f ** 2).sum()
    params = [head64[0].weight, head64[0].bias, head64[2].weight, head64[2].bias]
    grads = torch.autograd.grad(loss_sum * scale, [jets64] + params)
    return loss_sum.detach(), f.detach().reshape(-1), grads[0], grads[1:]


@pytest.mark.parametrize("residual", ["helmholtz", "t2d", "laplace"])
@pytest.mark.parametrize("C", [4, 8, 16, 32])
@pytest.mark.parametrize("dim", [2, 3])
def test_pde_head_step_matches_torch_autograd(cuda, dim, C, residual):
    from cosinesampler_b200 import jet
    if residual == "t2d" and dim == 3:
        pytest.skip("t2d is a 2D residual")
    gen = torch.Generator().manual_seed(7 * C + dim)
    J = 1 + 2 * dim
    for P in (1003, 4096):
        jets = torch.randn(J, C, P, generator=gen)
        jets[1:1 + dim] *= 3.0
        jets[1 + dim:] *= 9.0
        head32 = make_head(C, seed=4).to(cuda)
        head64 = make_head(C, seed=4, dtype=torch.float64)
        scale = 0.37 / P
        k2 = 9.8696
        loss_sum, gJets, grads, f = jet.pde_head_step(jets.to(cuda), head32, dim, residual, k2, scale, want_f=True)
        r_loss, r_f, r_gJets, r_grads = _torch_head_reference(jets.double(), head64, dim, residual, k2, scale)
        what = "head %dD C=%d P=%d %s " % (dim, C, P, residual)
        assert_close_scaled(f, r_f, what + "f")
        assert_close_scaled(loss_sum, r_loss, what + "loss_sum", rtol=1e-5)
        assert_close_scaled(gJets, r_gJets, what + "gJets", atol_scale=2e-6)
        for name, a, b in zip(("gW1", "gb1", "gw2", "gb2"), grads, r_grads):
            assert_close_scaled(a, b, what + name, rtol=2e-5, atol_scale=2e-5)
        # in place: the gradient overwrites the jets
        buf = jets.to(cuda)
        _, g2, _, _ = jet.pde_head_step(buf, head32, dim, residual, k2, scale, in_place=True)
        assert g2.data_ptr() == buf.data_ptr()
        assert torch.equal(g2, gJets)


def test_fused_pde_step_equals_dropin_training_step(cuda):
    """jet.fused_pde_step (3 launches per chunk, no autograd) == chain.training_step through the
    drop-in operator (14 launches + torch autograd): loss, cells.grad, head grads."""
    from cosinesampler_b200 import jet
    from cosinesampler_b200.chain import training_step
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    for dim, shape, kernel, residual, D in ((2, (4, 16, 64, 64), "cosine", "helmholtz", CosineSampler2d),
                                            (2, (4, 16, 64, 64), "cosine", "t2d", CosineSampler2d),
                                            (3, (4, 16, 16, 16, 16), "smooth-step", "laplace", CosineSampler3d)):
        gen = torch.Generator().manual_seed(11 + dim)
        P = 8192
        cells0 = torch.rand(shape, generator=gen)
        cols = [(torch.rand(P, 1, generator=gen) * 2 - 1).to(cuda) for _ in range(dim)]
        coords = torch.cat(cols, -1).contiguous()
        res = {}
        for mode in ("dropin", "fused", "fused_chunked", "autograd_head"):
            cells = torch.nn.Parameter(cells0.clone().to(cuda))
            head = make_head(shape[1], seed=3).to(cuda)
            if mode == "dropin":
                loss = training_step(lambda c, g: D.apply(c, g, "zeros", True, kernel, True), cells, cols, head,
                                     residual)
            elif mode == "autograd_head":
                S = jet.SamplerJet2d if dim == 2 else jet.SamplerJet3d
                jets = S.apply(cells, coords, "zeros", True, kernel, True)
                loss = jet.pde_head_loss(jets, head, dim, residual)
                loss.backward()
            else:
                loss = jet.fused_pde_step(cells, coords, head, residual, kernel=kernel,
                                          chunk=3000 if mode == "fused_chunked" else None)
            res[mode] = (loss.detach(), cells.grad, [p.grad for p in head.parameters()])
        for mode in ("fused", "fused_chunked", "autograd_head"):
            what = "%dD %s %s " % (dim, residual, mode)
            assert_close_scaled(res[mode][0], res["dropin"][0], what + "loss", rtol=1e-4)
            assert_close_scaled(res[mode][1], res["dropin"][1], what + "cells.grad", rtol=1e-4, atol_scale=2e-5)
            for a, b in zip(res[mode][2], res["dropin"][2]):
                assert_close_scaled(a, b, what + "head grad", rtol=1e-4, atol_scale=2e-5)


def test_fused_head_rejects_other_heads(cuda):
    from cosinesampler_b200 import jet
    jets = torch.rand(5, 16, 64, device=cuda)
    wide = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)).to(cuda)
    with pytest.raises(NotImplementedError):
        jet.pde_head_step(jets, wide, 2)
    # jet_mlp (torch ops) takes any Linear/Tanh stack
    u, first, second = jet.jet_mlp(wide, jets, 2)
    assert u.shape == (64, 1) and len(first) == 2 and len(second) == 2


def test_fused_pde_step_falls_back_for_other_heads(cuda):
    """A wider / deeper Linear-Tanh head takes the jet kernels + torch head path and still equals the
    drop-in chain."""
    from cosinesampler_b200 import jet
    from cosinesampler_b200.chain import training_step
    from cosine_sampler_2d import CosineSampler2d
    gen = torch.Generator().manual_seed(21)
    P, C = 4096, 8
    cells0 = torch.rand(3, C, 32, 32, generator=gen)
    cols = [(torch.rand(P, 1, generator=gen) * 2 - 1).to(cuda) for _ in range(2)]

    def make():
        torch.manual_seed(4)
        return torch.nn.Sequential(torch.nn.Linear(C, 24), torch.nn.Tanh(), torch.nn.Linear(24, 12), torch.nn.Tanh(),
                                   torch.nn.Linear(12, 1)).to(cuda)
    res = {}
    for mode in ("dropin", "fused"):
        cells = torch.nn.Parameter(cells0.clone().to(cuda))
        head = make()
        assert not jet.head_is_fusable(head, C)
        if mode == "dropin":
            loss = training_step(lambda c, g: CosineSampler2d.apply(c, g, "zeros", True, "cosine", True), cells, cols,
                                 head, "helmholtz")
        else:
            loss = jet.fused_pde_step(cells, torch.cat(cols, -1).contiguous(), head, "helmholtz", chunk=1500)
        res[mode] = (loss.detach(), cells.grad, [p.grad for p in head.parameters()])
    assert_close_scaled(res["fused"][0], res["dropin"][0], "fallback loss", rtol=1e-4)
    assert_close_scaled(res["fused"][1], res["dropin"][1], "fallback cells.grad", rtol=1e-4, atol_scale=2e-5)
    for a, b in zip(res["fused"][2], res["dropin"][2]):
        assert_close_scaled(a, b, "fallback head grad", rtol=1e-4, atol_scale=2e-5)


@pytest.mark.parametrize("dim", [2, 3])
def test_mixed_second_derivatives_match_nested_autograd(cuda, dim):
    """order 3 = order 2 + the mixed second derivatives (cu3d:836-856): u_ab from SamplerJet + jet_mlp equals
    nested autograd through the oracle sampler in fp64, and so do u, u_a, u_aa; the backward reaches the cells."""
    from cosinesampler_b200 import jet
    fn = grid_sample_2d if dim == 2 else grid_sample_3d
    S = jet.SamplerJet2d if dim == 2 else jet.SamplerJet3d
    gen = torch.Generator().manual_seed(60 + dim)
    shape = (4, 8, 20, 20) if dim == 2 else (3, 8, 10, 10, 10)
    P = 3000
    cells0 = torch.rand(shape, generator=gen)
    coords0 = safe_coords(P, dim, shape[2:][::-1], shape[0], True, gen).float()
    head32 = make_head(shape[1], seed=5).to(cuda)
    head64 = make_head(shape[1], seed=5, dtype=torch.float64)
    cells = cells0.to(cuda).requires_grad_(True)
    jets = S.apply(cells, coords0.to(cuda).contiguous(), "zeros", True, "cosine", True, 3)
    assert jets.shape[0] == jet.jet_count(dim, 3)
    u, u_a, u_aa, u_ab = jet.jet_mlp(head32, jets, dim, order=3)
    # fp64 oracle: nested autograd, mixed terms included
    c64 = cells0.double().requires_grad_(True)
    cols = [coords0[:, a:a + 1].double().requires_grad_(True) for a in range(dim)]
    grid = torch.cat(cols, -1).reshape((1,) * dim + (P, dim)).repeat((shape[0],) + (1,) * (dim + 1))
    val = fn(c64, grid, step="cosine", offset=True)
    ur = head64(val.sum(0).reshape(shape[1], -1).t())
    g1 = [torch.autograd.grad(ur.sum(), cols[a], create_graph=True)[0] for a in range(dim)]
    assert_close_scaled(u, ur, "u", rtol=1e-4, atol_scale=2e-5)
    for a in range(dim):
        assert_close_scaled(u_a[a], g1[a], "u_%d" % a, rtol=1e-4, atol_scale=2e-5)
        gaa = torch.autograd.grad(g1[a].sum(), cols[a], retain_graph=True)[0]
        assert_close_scaled(u_aa[a], gaa, "u_%d%d" % (a, a), rtol=1e-4, atol_scale=2e-5)
    for (a, b), v in u_ab.items():
        gab = torch.autograd.grad(g1[a].sum(), cols[b], retain_graph=True)[0]
        assert_close_scaled(v, gab, "u_%d%d" % (a, b), rtol=1e-4, atol_scale=2e-5)
    # a residual with a mixed term: its gradient w.r.t. the cells through the jets' backward
    f = sum(u_aa) + 0.7 * sum(u_ab.values()) + u
    loss = (f ** 2).mean()
    loss.backward()
    fr = ur
    for a in range(dim):
        fr = fr + torch.autograd.grad(g1[a].sum(), cols[a], create_graph=True)[0]
        for b in range(a + 1, dim):
            fr = fr + 0.7 * torch.autograd.grad(g1[a].sum(), cols[b], create_graph=True)[0]
    lr = (fr ** 2).mean()
    gr = torch.autograd.grad(lr, c64)[0]
    assert_close_scaled(loss, lr, "mixed loss", rtol=1e-4)
    assert_close_scaled(cells.grad, gr, "mixed d loss / d cells", rtol=1e-4, atol_scale=2e-5)
