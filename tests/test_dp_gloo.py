"""N>1 host logic on CPU: two gloo ranks, flat-buffer gradient all-reduce + fused Adam (through the
test-only CPU emulation of te_adam_ema).  Checks DDP semantics: every rank ends with the same
parameters == single-process Adam on the rank-averaged gradient; skipped tail groups stay untouched."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Setter:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_module():
    torch.manual_seed(5)
    m = torch.nn.Module()
    m.a = torch.nn.Linear(7, 5)
    m.to_rgb_like = torch.nn.Parameter(torch.randn(3))
    m.b = torch.nn.Linear(5, 2, bias=False)
    return m


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import emu
    emu.install(_Setter())
    from transeditor_b200.train_step import FlatAdam, FlatParams
    m = _make_module()
    flat = FlatParams(m, [lambda n: n == "to_rgb_like"])
    opt = FlatAdam(flat, 0.01, (0.0, 0.99))
    assert flat.group_end[0] % 4 == 0 and flat.group_end[1] == flat.numel
    for step in range(3):
        flat.clear_grads()
        g = torch.Generator().manual_seed(100 * step + rank)
        for _, p in flat.params:
            p.grad = torch.randn(p.shape, generator=g)       # what AccumulateGrad leaves behind
        flat.gather_grads()
        dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
        opt.step(1 if step == 1 else 2, grad_scale=1.0 / world)  # step 1 skips the tail group
    out[rank] = flat.data.clone()
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_adam_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.equal(out[0], out[1])
    # single-process reference with torch.optim.Adam on averaged gradients
    m = _make_module()
    main = [p for n, p in m.named_parameters() if n != "to_rgb_like"]
    tail = [m.to_rgb_like]
    o_main = torch.optim.Adam(main, lr=0.01, betas=(0.0, 0.99))
    o_tail = torch.optim.Adam(tail, lr=0.01, betas=(0.0, 0.99))
    for step in range(3):
        grads = {}
        for rank in range(world):
            g = torch.Generator().manual_seed(100 * step + rank)
            # same ordering as FlatParams: main group first, tail group last
            for n, p in [(n, p) for n, p in m.named_parameters() if n != "to_rgb_like"] + [("to_rgb_like", m.to_rgb_like)]:
                grads[n] = grads.get(n, 0) + torch.randn(p.shape, generator=g) / world
        for n, p in m.named_parameters():
            p.grad = grads[n]
        o_main.step()
        if step != 1:
            o_tail.step()
    from transeditor_b200.train_step import FlatParams
    ref = FlatParams(_make_module(), [lambda n: n == "to_rgb_like"])
    for n, p in m.named_parameters():
        o = ref.offsets[n]
        got = out[0][o:o + p.numel()].view(p.shape)
        assert torch.allclose(got, p.detach(), atol=1e-6), n
