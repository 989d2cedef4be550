"""te_attn_stack_fwd / te_attn_stack_bwd (the interaction network in one launch) against the float64 CPU
restatement of AttentionBlock / Attention (model_spatial_query.py:883-936) in oracle/te_oracle.py
(`interaction_stack` over `attention_block`), which tests/test_oracle_golden.py pins to the reference's own classes."""
import os

import pytest
import torch

from oracle import te_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
LR = 0.01
FIELDS = ("w_proj", "b_proj", "w_q", "b_q", "w_k", "b_k", "w_v", "b_v", "w_o", "b_o", "w_m1", "b_m1", "w_m2", "b_m2")


def _blocks(n_blocks, first_dim, seed, dtype=torch.float32, device="cpu"):
    g = torch.Generator().manual_seed(seed)

    def w(o, i):
        return (torch.randn(o, i, generator=g, dtype=torch.float64) / LR).to(dtype).to(device)

    def b(o):
        return (torch.randn(o, generator=g, dtype=torch.float64) * 0.3 / LR).to(dtype).to(device)

    out = []
    for i in range(n_blocks):
        d = first_dim if i == 0 else 512
        blk = {"in_dim": d, "param_dim": d,
               "w_proj": w(512, d) if d != 512 else None, "b_proj": b(512) if d != 512 else None,
               "w_q": w(128, d), "b_q": b(128), "w_k": w(128, d), "b_k": b(128), "w_v": w(128, d), "b_v": b(128),
               "w_o": w(512, 128), "b_o": b(512), "w_m1": w(512, 512), "b_m1": b(512), "w_m2": w(512, 512),
               "b_m2": b(512)}
        out.append(blk)
    return out


def _inputs(batch, first_dim, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(batch, 16, first_dim, generator=g, dtype=torch.float64)
    p0 = torch.randn(batch, 16, first_dim, generator=g, dtype=torch.float64)
    p = torch.randn(batch, 16, 512, generator=g, dtype=torch.float64)
    return x0, p0, p


def _to(blocks, dtype, device, grad=False):
    out = []
    for blk in blocks:
        new = {}
        for k, v in blk.items():
            if torch.is_tensor(v):
                v = v.detach().to(dtype).to(device)
                if grad:
                    v.requires_grad_(True)
            new[k] = v
        out.append(new)
    return out


def _rel(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item() / max(1e-30, b.double().abs().max().item())


@pytest.mark.parametrize("tf32", [False, True])
@pytest.mark.parametrize("batch,n_blocks,first_dim", [(3, 8, 528), (16, 8, 528), (1, 1, 528), (2, 2, 512), (5, 3, 64),
                                                      (35, 2, 528), (20, 2, 528)])
def test_forward_matches_float64_restatement(batch, n_blocks, first_dim, tf32):
    from transeditor_b200 import op
    b64 = _blocks(n_blocks, first_dim, seed=n_blocks, dtype=torch.float64)
    x0, p0, p = _inputs(batch, first_dim, seed=batch)
    ref = O.interaction_stack(x0, p0, p, b64)
    bd = _to(b64, torch.float32, DEV)
    pd = p.float().to(DEV) if n_blocks > 1 else None
    assert op.attn_stack_supported(x0.float().to(DEV), p0.float().to(DEV), pd, bd)
    with torch.no_grad():
        y = op.attn_stack(x0.float().to(DEV), p0.float().to(DEV), pd, bd, LR, tf32=tf32)
    assert y.shape == (batch, 16, 512) and y.dtype == torch.float32
    # f32 in/out with 3xTF32 products: a few f32 ulps per layer; single-pass TF32: 2^-11 per operand
    assert _rel(y, ref) < (1e-2 if tf32 else 2e-5)


@pytest.mark.parametrize("tf32", [False, True])
@pytest.mark.parametrize("batch,n_blocks,first_dim", [(3, 8, 528), (16, 8, 528), (2, 1, 528), (2, 2, 512), (4, 3, 64)])
def test_backward_matches_float64_autograd(batch, n_blocks, first_dim, tf32):
    from transeditor_b200 import op
    b64 = _to(_blocks(n_blocks, first_dim, seed=10 + n_blocks, dtype=torch.float64), torch.float64, "cpu", grad=True)
    x0, p0, p = [t.requires_grad_(True) for t in _inputs(batch, first_dim, seed=20 + batch)]
    gy = torch.randn(batch, 16, 512, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    ref = O.interaction_stack(x0, p0, p, b64)
    leaves = [x0, p0] + ([p] if n_blocks > 1 else []) + [blk[f] for blk in b64 for f in FIELDS if blk[f] is not None]
    gref = torch.autograd.grad(ref, leaves, gy)

    bd = _to(b64, torch.float32, DEV, grad=True)
    xd, p0d, pd = [t.detach().float().to(DEV).requires_grad_(True) for t in (x0, p0, p)]
    y = op.attn_stack(xd, p0d, pd if n_blocks > 1 else None, bd, LR, tf32=tf32)
    dleaves = [xd, p0d] + ([pd] if n_blocks > 1 else []) + [blk[f] for blk in bd for f in FIELDS if blk[f] is not None]
    got = torch.autograd.grad(y, dleaves, gy.float().to(DEV))
    names = ["x0", "p0"] + (["p"] if n_blocks > 1 else []) + ["%d.%s" % (i, f) for i, blk in enumerate(b64)
                                                                 for f in FIELDS if blk[f] is not None]
    # b_k's true gradient is zero (a key bias shifts every logit of a softmax row equally), so errors are
    # measured against the largest gradient of the same kind as well
    kind = lambda n: n.split(".")[-1][:2] if "." in n else n  # noqa: E731  ("w_", "b_" or an input's name)
    floor = {kind(n): 0.0 for n in names}
    for name, b in zip(names, gref):
        floor[kind(name)] = max(floor[kind(name)], b.abs().max().item())
    for name, a, b in zip(names, got, gref):
        assert a.shape == b.shape, name
        err = (a.double().cpu() - b).abs().max().item()
        tol = 3e-2 if tf32 else 5e-5
        assert err < tol * max(b.abs().max().item(), 1e-2 * floor[kind(name)]), "%s: err %g (max %g)" % (
            name, err, b.abs().max().item())


def test_second_order_route_is_differentiable():
    """create_graph=True re-expresses the backward with differentiable ops: d/dp of |dy/dp0|^2 must match the
    all-torch restatement (the optional spatial path regulariser's derivative, train_spatial_query.py:252-277)."""
    from transeditor_b200 import op
    b64 = _blocks(2, 528, seed=3, dtype=torch.float64)
    x0, p0, p = _inputs(2, 528, seed=4)

    def penalty(fn, x0, p0, p, blocks):
        p0 = p0.requires_grad_(True)
        p = p.requires_grad_(True)
        y = fn(x0, p0, p, blocks, LR)
        (g,) = torch.autograd.grad(y.square().sum(), p0, create_graph=True)
        (gg,) = torch.autograd.grad(g.square().sum(), p)
        return gg

    ref = penalty(lambda a, b_, c, blks, lr: O.interaction_stack(a, b_, c, blks), x0, p0, p, b64)
    got = penalty(op.attn_stack, x0.float().to(DEV), p0.float().to(DEV), p.float().to(DEV),
                  _to(b64, torch.float32, DEV))
    assert _rel(got, ref) < 1e-3


def test_generator_latent_same_with_and_without_the_fused_stack():
    import model_spatial_query as M
    torch.manual_seed(0)
    g = M.Generator(32, 512, 512, 8, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to(DEV)
    z, p = torch.randn(4, 512, 16, device=DEV), torch.randn(4, 512, 16, device=DEV)
    params = [q for n, q in g.named_parameters() if n.startswith(("interact", "style_mapping", "spatial_mapping"))]

    def run():
        lat = g(z, p, return_only_style_latent=True)
        grads = torch.autograd.grad(lat.square().sum(), params, allow_unused=True)
        return lat.detach(), grads

    lat1, g1 = run()
    os.environ["TE_ATTN_STACK"] = "0"
    try:
        lat0, g0 = run()
    finally:
        del os.environ["TE_ATTN_STACK"]
    assert _rel(lat1, lat0) < 1e-5
    # several of these gradients are mathematically zero (key biases under a softmax): absolute scale
    scale = max(b.abs().max().item() for b in g0 if b is not None)
    for a, b in zip(g1, g0):
        assert (a is None) == (b is None)
        if a is not None:
            assert (a - b).abs().max().item() < 2e-4 * max(b.abs().max().item(), 1e-2 * scale)


def test_replays_from_a_cuda_graph():
    """Forward + backward captured once and replayed, the way train_step.Trainer runs a phase."""
    from transeditor_b200 import op
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        bd = _to(_blocks(8, 528, seed=1), torch.float32, DEV, grad=True)
        x0, p0, p = [t.float().to(DEV).requires_grad_(True) for t in _inputs(4, 528, seed=2)]
        params = [blk[f] for blk in bd for f in FIELDS if blk[f] is not None]

        def step():
            y = op.attn_stack(x0, p0, p, bd, LR)
            return (y,) + torch.autograd.grad(y.square().sum(), [x0, p0, p] + params)

        for _ in range(3):
            eager = [t.clone() for t in step()]
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    for a, b in zip(out, eager):
        assert _rel(a, b) < 1e-6


def test_rejects_bad_tables():
    from transeditor_b200 import lib
    bd = _to(_blocks(2, 528, seed=1), torch.float32, DEV)
    x0, p0, p = [t.float().to(DEV) for t in _inputs(2, 528, seed=2)]
    y = torch.empty(2, 16, 512, device=DEV)
    bad = [dict(bd[0]), dict(bd[1])]
    bad[1]["in_dim"] = 528
    with pytest.raises(RuntimeError, match="512 wide"):
        lib.attn_stack_fwd(y, x0, p0, p, bad, 2, LR, False, None)
    bad = [dict(bd[0]), dict(bd[1])]
    bad[0]["w_proj"] = None
    with pytest.raises(RuntimeError, match="w_proj"):
        lib.attn_stack_fwd(y, x0, p0, p, bad, 2, LR, False, None)
    with pytest.raises(RuntimeError, match="n_blocks"):
        lib.attn_stack_fwd(y, x0, p0, p, bd * 5, 2, LR, False, None)


@pytest.mark.parametrize("create_graph", [False, True])
def test_partial_derivatives_when_arguments_depend_on_each_other(create_graph):
    """Generator.forward passes p AND p0 = cat(p, eye): both backward routes (kernel / differentiable re-expression)
    must return PARTIAL derivatives per argument, or the p0 path is counted twice."""
    from transeditor_b200 import op
    bd = _to(_blocks(3, 528, seed=5), torch.float32, DEV, grad=True)
    g = torch.Generator().manual_seed(6)
    q = torch.randn(2, 16, 512, generator=g).to(DEV).requires_grad_(True)
    s = torch.randn(2, 16, 512, generator=g).to(DEV).requires_grad_(True)
    eye = torch.eye(16, device=DEV).repeat(2, 1, 1)
    gy = torch.randn(2, 16, 512, generator=g).to(DEV)

    def run(fn):
        y = fn(torch.cat([s, eye], 2), torch.cat([q, eye], 2), q, bd, LR)
        return torch.autograd.grad(y, [q, s], gy, create_graph=create_graph)

    for a, b in zip(run(op.attn_stack), run(lambda a_, b_, c, blks, lr: O.interaction_stack(a_, b_, c, blks))):
        assert _rel(a, b) < 1e-4
