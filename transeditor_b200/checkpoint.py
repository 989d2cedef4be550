"""Checkpoints in the reference's own format.

`train_spatial_query.py:361-371` saves `{'g', 'd', 'g_ema', 'g_optim', 'd_optim'}`: three module state dicts and two
`torch.optim.Adam.state_dict()`s, and `:475-492` loads them back.  Our trainer keeps parameters and Adam moments
in flat buffers (train_step.FlatParams / FlatAdam), so the optimiser state is converted to and from Adam's
per-parameter layout here: parameter index = position in `module.parameters()`, per-parameter `step`,
`exp_avg`, `exp_avg_sq`.  A file written by either side resumes on the other.
"""
import torch


def adam_state_dict(optim, module):
    """FlatAdam -> the dict torch.optim.Adam(module.parameters(), lr, betas).state_dict() would hold.  Parameters
    whose group was never stepped (Adam skips grad=None) have no state entry, exactly like torch's."""
    flat = optim.flat
    steps = [int(s.item()) for s in optim.steps]
    state = {}
    names = [n for n, _ in module.named_parameters()]
    by_name = dict(flat.params)
    for idx, name in enumerate(names):
        off = flat.offsets[name]
        gi = next(g for g, end in enumerate(flat.group_end) if off < end)
        if steps[gi] == 0:
            continue
        p = by_name[name]
        n = p.numel()
        state[idx] = {"step": torch.tensor(float(steps[gi])),
                      "exp_avg": optim.m[off:off + n].view(p.shape).clone(),
                      "exp_avg_sq": optim.v[off:off + n].view(p.shape).clone()}
    group = {"lr": optim.lr, "betas": (optim.betas[0], optim.betas[1]), "eps": optim.eps, "weight_decay": 0,
             "amsgrad": False, "params": list(range(len(names)))}
    return {"state": state, "param_groups": [group]}


def load_adam_state_dict(optim, module, sd):
    """torch.optim.Adam state dict (reference checkpoint) -> FlatAdam moments, step counters and hyper-parameters."""
    flat = optim.flat
    names = [n for n, _ in module.named_parameters()]
    by_name = dict(flat.params)
    order = []
    for g in sd["param_groups"]:
        order += list(g["params"])
    if len(order) != len(names):
        raise ValueError("optimizer state has %d parameters, the module has %d" % (len(order), len(names)))
    optim.m.zero_()
    optim.v.zero_()
    group_step = [None] * len(flat.group_end)
    for pos, idx in enumerate(order):
        st = sd["state"].get(idx, sd["state"].get(str(idx)))
        name = names[pos]
        off = flat.offsets[name]
        gi = next(g for g, end in enumerate(flat.group_end) if off < end)
        if st is None:
            step = 0
        else:
            p = by_name[name]
            n = p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("optimizer state of %s has shape %s, expected %s" %
                                 (name, tuple(st["exp_avg"].shape), tuple(p.shape)))
            optim.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
            optim.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            step = int(float(st["step"]))
        if group_step[gi] is None:
            group_step[gi] = step
        elif group_step[gi] != step:
            raise ValueError("parameters of update group %d disagree on the Adam step count (%d vs %d): %s" %
                             (gi, group_step[gi], step, name))
    for gi, step in enumerate(group_step):
        optim.steps[gi].fill_(step or 0)
    g0 = sd["param_groups"][0]
    optim.lr = float(g0["lr"])
    optim.betas = (float(g0["betas"][0]), float(g0["betas"][1]))
    optim.eps = float(g0["eps"])


def checkpoint_dict(trainer):
    """The dict `train_spatial_query.py:362-370` saves."""
    return {"g": trainer.generator.state_dict(), "d": trainer.discriminator.state_dict(),
            "g_ema": trainer.g_ema.state_dict(),
            "g_optim": adam_state_dict(trainer.g_optim, trainer.generator),
            "d_optim": adam_state_dict(trainer.d_optim, trainer.discriminator)}


def save_checkpoint(trainer, path):
    torch.save(checkpoint_dict(trainer), path)


def iteration_from_name(path):
    """The reference derives the start iteration from the file name (`790000.pt` -> 790000,
    train_spatial_query.py:478-483); None when the name is not a number."""
    import os
    stem = os.path.splitext(os.path.basename(str(path)))[0]
    try:
        return int(stem)
    except ValueError:
        return None


def load_checkpoint(trainer, path_or_dict, strict=True, iteration=None):
    """Resume from a reference-format checkpoint (ours or the authors').  Parameters are copied IN PLACE into the
    flat buffers (captured CUDA graphs stay valid).  Files without optimiser state (the released inference
    checkpoints hold only 'g_ema') load the networks and leave the optimisers untouched.

    The file is read with `weights_only=True` (tensors, dicts and numbers only: a third-party checkpoint cannot run
    pickle code).  `trainer.iteration` — which drives the d_reg / g_reg cadence — is set from `iteration`, else from
    the file name like the reference does, else left alone.  `mean_path_length` is not part of the reference's
    checkpoint and restarts at zero there too."""
    if isinstance(path_or_dict, dict):
        ckpt = path_or_dict
    else:
        ckpt = torch.load(path_or_dict, map_location="cpu", weights_only=True)
        if iteration is None:
            iteration = iteration_from_name(path_or_dict)
    if iteration is not None:
        trainer.iteration = int(iteration)
    if "g" in ckpt:
        trainer.generator.load_state_dict(ckpt["g"], strict=strict)
    if "d" in ckpt:
        trainer.discriminator.load_state_dict(ckpt["d"], strict=strict)
    if "g_ema" in ckpt:
        trainer.g_ema.load_state_dict(ckpt["g_ema"], strict=strict)
        if "g" not in ckpt:
            trainer.generator.load_state_dict(ckpt["g_ema"], strict=strict)
    if "g_optim" in ckpt:
        load_adam_state_dict(trainer.g_optim, trainer.generator, ckpt["g_optim"])
    if "d_optim" in ckpt:
        load_adam_state_dict(trainer.d_optim, trainer.discriminator, ckpt["d_optim"])
    # captured graphs read the bf16 weight copies of the repack cache without looking at version counters
    packs = getattr(trainer, "_packs", None)
    if packs is not None:
        packs.refresh()
    return ckpt
