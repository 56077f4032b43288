"""The C-ABI entry points are enqueue-only (include/cmcd_b200.h): a whole train step -- forward bridge, adjoint,
partial-gradient reduce -- can be captured into a CUDA graph and replayed, which is how a jitted caller (XLA, or a
torch.cuda.CUDAGraph around the training loop) drives them."""
import pytest
import torch

from helpers import oracle_problem, product_problem, seeds_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["C_manygmm_dds_small", "B_funnel", "D_lgcp"])
def test_train_step_replays_from_a_cuda_graph(name):
    from cmcd_b200 import _lib, mcdboundingmachine as PM
    c, lp, dim, pf, unf, fixed = oracle_problem(name)
    _, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    gl = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, **kw))
    seeds_a = torch.from_numpy(seeds_for(c["N"], seed=1)).cuda()
    seeds_b = torch.from_numpy(seeds_for(c["N"], seed=2)).cuda()
    g_a, (l_a, z_a) = gl(seeds_a, pf_p, unf_p, fixed_p, target)           # eager references (also warms everything up)
    g_b, (l_b, z_b) = gl(seeds_b, pf_p, unf_p, fixed_p, target)
    static_seeds = seeds_a.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gl(static_seeds, pf_p, unf_p, fixed_p, target)                    # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_s, (l_s, z_s) = gl(static_seeds, pf_p, unf_p, fixed_p, target)
    for seeds, (g_ref, l_ref, z_ref) in ((seeds_b, (g_b, l_b, z_b)), (seeds_a, (g_a, l_a, z_a))):
        static_seeds.copy_(seeds)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(l_s, l_ref) and torch.equal(z_s, z_ref)       # forward is bitwise reproducible
        scale = g_ref.abs().max().item()
        assert (g_s - g_ref).abs().max().item() <= 5e-5 * scale           # gradient: atomics change the fp32 summation order
