"""Reference-format checkpoints (train_spatial_query.py:361-371, 475-492): FlatAdam <-> torch.optim.Adam state."""
import io

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.usefixtures("cpu_emulation")


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.a = nn.Linear(5, 7)
        self.tail = nn.Parameter(torch.randn(3))       # like to_rgb biases: no gradient in some phases
        self.b = nn.Linear(7, 2, bias=False)


def _grads(m, seed):
    g = torch.Generator().manual_seed(seed)
    return {n: torch.randn(p.shape, generator=g) for n, p in m.named_parameters()}


def _flat_step(flat, opt, grads, n_groups):
    flat.clear_grads()
    for n, p in flat.params:
        if n_groups == 2 or n != "tail":
            p.grad = grads[n].clone()
    flat.gather_grads()
    opt.step(n_groups)


def test_flat_adam_state_round_trips_through_torch_adam():
    from transeditor_b200.checkpoint import adam_state_dict, load_adam_state_dict
    from transeditor_b200.train_step import FlatAdam, FlatParams
    ours, ref = _Net(), _Net()
    flat = FlatParams(ours, [lambda n: n == "tail"])
    opt = FlatAdam(flat, 0.01, (0.0, 0.99))
    t_main = torch.optim.Adam([p for n, p in ref.named_parameters()], lr=0.01, betas=(0.0, 0.99))
    for step in range(3):
        gr = _grads(ours, step)
        _flat_step(flat, opt, gr, 2)
        for n, p in ref.named_parameters():
            p.grad = gr[n].clone()
        t_main.step()
    sd = adam_state_dict(opt, ours)
    tsd = t_main.state_dict()
    assert sorted(sd["state"]) == sorted(tsd["state"]) and sd["param_groups"][0]["params"] == tsd["param_groups"][0]["params"]
    for i in tsd["state"]:
        assert float(sd["state"][i]["step"]) == float(tsd["state"][i]["step"]) == 3
        assert torch.allclose(sd["state"][i]["exp_avg"], tsd["state"][i]["exp_avg"], atol=1e-6)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], tsd["state"][i]["exp_avg_sq"], atol=1e-6)
    # torch's Adam accepts our dict (through a save/load cycle, like a resumed reference run) ...
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    fresh = torch.optim.Adam([p for _, p in _Net().named_parameters()], lr=0.5)
    fresh.load_state_dict(torch.load(buf))
    assert fresh.param_groups[0]["lr"] == 0.01
    # ... and a torch Adam state loads into a fresh FlatAdam, after which both take the same next step
    ours2 = _Net()
    ours2.load_state_dict(ref.state_dict())
    flat2 = FlatParams(ours2, [lambda n: n == "tail"])
    opt2 = FlatAdam(flat2, 0.5, (0.5, 0.5))
    load_adam_state_dict(opt2, ours2, tsd)
    assert opt2.lr == 0.01 and opt2.betas == (0.0, 0.99) and [int(s) for s in opt2.steps] == [3, 3]
    gr = _grads(ours, 99)
    _flat_step(flat2, opt2, gr, 2)
    for n, p in ref.named_parameters():
        p.grad = gr[n].clone()
    t_main.step()
    for (n, p), (_, q) in zip(ours2.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p, q, atol=1e-6), n


def test_unstepped_tail_group_has_no_state():
    from transeditor_b200.checkpoint import adam_state_dict, load_adam_state_dict
    from transeditor_b200.train_step import FlatAdam, FlatParams
    net = _Net()
    flat = FlatParams(net, [lambda n: n == "tail"])
    opt = FlatAdam(flat, 0.01, (0.0, 0.99))
    _flat_step(flat, opt, _grads(net, 0), 1)           # the tail group got no gradient
    sd = adam_state_dict(opt, net)
    names = [n for n, _ in net.named_parameters()]
    assert names.index("tail") not in sd["state"] and len(sd["state"]) == len(names) - 1
    opt2 = FlatAdam(FlatParams(_Net(), [lambda n: n == "tail"]), 0.01, (0.0, 0.99))
    load_adam_state_dict(opt2, net, sd)
    assert [int(s) for s in opt2.steps] == [1, 0]
