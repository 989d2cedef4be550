"""N>1 host logic on CPU: two gloo ranks, flat-buffer gradient all-reduce + fused Adam (through the
test-only CPU emulation of te_adam_ema).  Checks DDP semantics: every rank ends with the same
parameters == single-process Adam on the rank-averaged gradient; skipped tail groups stay untouched."""
import os
import socket

import pytest

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


def _bucket_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transeditor_b200.train_step import FlatParams, GradBuckets
    torch.manual_seed(5)
    m = torch.nn.Sequential(torch.nn.Linear(9, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                            torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.ones(5))           # never receives a gradient (the noise strengths' case)
    m.register_parameter("unused", unused)
    flat = FlatParams(m, [lambda n: n == "unused"])
    buckets = GradBuckets(flat, world, bucket_mb=100 * 4 / (1 << 20))   # ~100 floats per bucket -> several buckets
    assert len(buckets.buckets) >= 3
    assert buckets.buckets[0][0] == 0 and buckets.buckets[-1][1] == flat.numel
    for step in range(2):
        x = torch.randn(6, 9, generator=torch.Generator().manual_seed(10 * step + rank))
        buckets.begin()
        m(x).square().sum().backward()
        buckets.finish()
        assert all(p.grad is None for _, p in flat.params)
    out[rank] = flat.grad.clone()
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_sums_the_ranks_gradients():
    """GradBuckets (hooks + per-bucket all-reduce during backward) == the plain sum of the ranks' gradients, zeros
    for parameters without a gradient."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.equal(out[0], out[1])
    torch.manual_seed(5)
    m = torch.nn.Sequential(torch.nn.Linear(9, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                            torch.nn.Linear(16, 4))
    m.register_parameter("unused", torch.nn.Parameter(torch.ones(5)))
    from transeditor_b200.train_step import FlatParams
    ref = FlatParams(m, [lambda n: n == "unused"])
    want = torch.zeros_like(ref.grad)
    for rank in range(world):
        x = torch.randn(6, 9, generator=torch.Generator().manual_seed(10 + rank))
        for _, p in ref.params:
            p.grad = None
        m(x).square().sum().backward()
        for n, p in ref.params:
            if p.grad is not None:
                o = ref.offsets[n]
                want[o:o + p.numel()] += p.grad.reshape(-1)
    assert torch.allclose(out[0], want, atol=1e-6)
    o = ref.offsets["unused"]
    assert out[0][o:o + 5].abs().sum().item() == 0


def _loss_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transeditor_b200.train_step import Trainer
    tr = Trainer.__new__(Trainer)      # only the reduction helpers are exercised: no models, no GPU
    tr.world, tr.rank = world, rank
    tr.losses = {"d": torch.tensor(1.0 + rank), "g": torch.tensor(10.0 * (rank + 1)), "r1": torch.tensor(0.5)}
    tr.mean_path_length = torch.tensor(2.0 + 4 * rank)
    keys, vec = tr.reduced_losses()
    out[rank] = (keys, vec.tolist(), float(tr.mean_path_length_avg()))
    dist.destroy_process_group()


def test_loss_dict_and_path_length_reduction_match_the_reference_semantics():
    """reduce_loss_dict (utils/distributed.py:102-124): sorted keys, reduce to rank 0, divided by the world size THERE
    (the other ranks keep what dist.reduce leaves them); reduce_sum(mean_path_length) / world on every rank."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_loss_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    keys, vec0, mpl0 = out[0]
    assert keys == ["d", "g", "r1"]
    assert vec0 == pytest.approx([1.5, 15.0, 0.5])
    assert mpl0 == pytest.approx(4.0) and out[1][2] == pytest.approx(4.0)
