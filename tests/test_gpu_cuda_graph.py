"""The operator chain has no host synchronisation (the reference has 8-19 `.item()` calls per step,
SURVEY 3.5), so a whole fwd -> triple-bwd step can be captured in a CUDA graph and replayed."""
import pytest
import torch

from util import assert_close_scaled

pytestmark = pytest.mark.gpu


def test_whole_pde_step_is_cuda_graph_capturable(cuda):
    from cosine_sampler_2d import CosineSampler2d
    from cosinesampler_b200 import chain
    gen = torch.Generator().manual_seed(5)
    cells = torch.nn.Parameter(torch.rand(4, 16, 24, 24, generator=gen).to(cuda))
    xy = (torch.rand(4096, 2, generator=gen) * 1.9 - 0.95).to(cuda)
    head = chain.make_head(16, seed=1, device=cuda)
    params = [cells] + list(head.parameters())
    sampler = lambda c, g: CosineSampler2d.apply(c, g, "zeros", True, "cosine", True)
    static_xy = xy.clone()

    def step():
        cols = [static_xy[:, a:a + 1].detach().requires_grad_(True) for a in range(2)]
        loss = chain.pde_loss(sampler, cells, cols, head, "helmholtz")
        grads = torch.autograd.grad(loss, params)
        return loss, grads

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()

    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_loss, g_grads = step()
    # new data in the static buffers, then replay
    with torch.no_grad():
        cells.copy_(torch.rand(cells.shape, generator=gen).to(cuda))
        static_xy.copy_((torch.rand(4096, 2, generator=gen) * 1.9 - 0.95).to(cuda))
    graph.replay()
    torch.cuda.synchronize()
    e_loss, e_grads = step()
    assert_close_scaled(g_loss, e_loss, "loss: graph replay vs eager", rtol=1e-5, atol_scale=1e-6)
    for a, b in zip(g_grads, e_grads):
        assert_close_scaled(a, b, "gradient: graph replay vs eager", rtol=1e-4, atol_scale=1e-5)


def test_fused_pde_step_is_cuda_graph_capturable(cuda):
    """The fused jet step is three launches and a few memsets / transposes per chunk, nothing
    else: capture and replay with new cells and points."""
    from cosinesampler_b200 import chain, jet
    gen = torch.Generator().manual_seed(6)
    cells = torch.nn.Parameter(torch.rand(4, 16, 24, 24, generator=gen).to(cuda))
    static_xy = (torch.rand(4096, 2, generator=gen) * 1.9 - 0.95).to(cuda)
    head = chain.make_head(16, seed=1, device=cuda)
    params = [cells] + list(head.parameters())

    def step():
        for p in params:
            p.grad = None
        loss = jet.fused_pde_step(cells, static_xy, head, "helmholtz", kernel="cosine", chunk=1500)
        return loss, [p.grad for p in params]

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_loss, g_grads = step()
    with torch.no_grad():
        cells.copy_(torch.rand(cells.shape, generator=gen).to(cuda))
        static_xy.copy_((torch.rand(4096, 2, generator=gen) * 1.9 - 0.95).to(cuda))
    graph.replay()
    torch.cuda.synchronize()
    g_loss = g_loss.clone()
    g_grads = [g.clone() for g in g_grads]
    e_loss, e_grads = step()
    assert_close_scaled(g_loss, e_loss, "fused loss: graph replay vs eager", rtol=1e-5, atol_scale=1e-6)
    for a, b in zip(g_grads, e_grads):
        assert_close_scaled(a, b, "fused gradient: graph replay vs eager", rtol=1e-4, atol_scale=1e-5)
