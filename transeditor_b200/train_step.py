"""Data-parallel G+D train step (the loop body of train_spatial_query.py:166-306) as an engine.

One process per GPU.  What the reference does with DistributedDataParallel + torch.optim.Adam +
a Python EMA loop, this does with

  * parameters, gradients, Adam moments and the EMA copy living in FLAT f32 buffers (one per
    model), so the gradient exchange is ONE NCCL all-reduce per optimiser step (G 172 MB,
    D 115 MB) and the optimiser + EMA is ONE te_adam_ema launch instead of ~650 tiny kernels;
  * the same losses, lazy-regularisation cadence and learning-rate / beta corrections
    (train_spatial_query.py:69-105,196-250,461-473).

Sharding: batch only (replicated weights, per-rank batch a multiple of 4 so the minibatch-stddev
groups match DDP semantics, model_spatial_query.py:845-852).  The only data-path collective is
the gradient sum (averaged by world size like DDP).
"""
import math
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import lib, tc
from .model import Discriminator, Generator


class TrainConfig:
    """Defaults = train_spatial_query.py:379-415."""

    def __init__(self, **kw):
        self.size = 256
        self.batch = 16
        self.latent = 512
        self.para_num = 16
        self.channel_multiplier = 2
        self.num_trans = 8
        self.pixel_norm_op_dim = 1
        self.inject_noise = False
        self.lr = 0.002
        self.r1 = 10.0
        self.path_regularize = 2.0
        self.path_batch_shrink = 2
        self.spatial_regu = False            # --spatial_regu (:405): second path regulariser, on the spatial code
        self.spatial_path_regularize = 2.0   # :388
        self.regu_space = "p+"               # --regu_sapce (:406): "p" = raw spatial code, "p+" = mapped code
        self.d_reg_every = 16
        self.g_reg_every = 4
        self.ema_decay = 0.5 ** (32 / (10 * 1000))  # :157
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown TrainConfig field %r" % k)
            setattr(self, k, v)

    @property
    def token(self):
        return 2 * (int(math.log2(self.size)) - 1)  # :432


class FlatParams:
    """Re-homes a module's parameters into one contiguous buffer (parameters become views) with a
    matching flat gradient buffer.  `tail_groups` is a list of predicates on the parameter name;
    matching parameters are placed, group by group, after everything else so that an optimiser
    step can leave them out (parameters the reference's Adam skips because their .grad is None)."""

    def __init__(self, module, tail_groups=()):
        named = list(module.named_parameters())
        buckets = [[] for _ in range(len(tail_groups) + 1)]
        for name, p in named:
            for gi, pred in enumerate(tail_groups):
                if pred(name):
                    buckets[gi + 1].append((name, p))
                    break
            else:
                buckets[0].append((name, p))
        ordered = [x for b in buckets for x in b]
        # every segment starts on a 16-byte boundary so the vectorised optimiser can run on ranges
        self.offsets, off = {}, 0
        self.group_end = []
        for b in buckets:
            for name, p in b:
                self.offsets[name] = off
                off += (p.numel() + 3) // 4 * 4
            self.group_end.append(off)
        self.numel = off
        dev, dt = ordered[0][1].device, ordered[0][1].dtype
        self.data = torch.zeros(off, dtype=dt, device=dev)
        self.grad = torch.zeros(off, dtype=dt, device=dev)
        self.params = []
        for name, p in ordered:
            o, n = self.offsets[name], p.numel()
            self.data[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.data[o:o + n].view(p.shape)
            p.grad = None
            self.params.append((name, p))

    def rebind_grads(self):
        for name, p in self.params:
            o, n = self.offsets[name], p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad[o:o + n].data_ptr():
                p.grad = self.grad[o:o + n].view(p.shape)

    def clear_grads(self):
        """Detach the parameters from the flat gradient buffer before a backward pass.  With .grad = None
        autograd's AccumulateGrad just keeps the incoming gradient tensor (no `grad += g` kernel per
        parameter: ~250 tiny launches per generator backward); gather_grads() then moves everything into
        the flat buffer with one multi-tensor copy."""
        for _, p in self.params:
            p.grad = None

    def gather_grads(self):
        self.grad.zero_()
        views, grads = [], []
        for name, p in self.params:
            if p.grad is not None:
                o = self.offsets[name]
                views.append(self.grad[o:o + p.numel()].view(p.shape))
                grads.append(p.grad)
        if grads:
            torch._foreach_copy_(views, grads)
        for _, p in self.params:
            p.grad = None


class GradBuckets:
    """DDP-style overlapped gradient exchange for a FlatParams (what torch's DistributedDataParallel does for the
    reference, train_spatial_query.py:495-509): the flat buffer is cut into contiguous buckets of ~`bucket_mb`; a
    post-accumulate-grad hook on every parameter counts arrivals, and as soon as a bucket's parameters all have their
    gradients these are copied into the bucket's slice of the flat buffer and ONE all-reduce of that slice is issued
    on a side stream — while the backward pass keeps running on the computing stream.  Backward produces the LAST
    layers' gradients first, so by the time it reaches the first layers most of the buffer is already reduced; only
    the final bucket's all-reduce is exposed.  `finish()` flushes what is left (parameters that received no gradient
    stay zero, like the flat gather they replace) and joins the side stream.  Everything is stream-ordered (events,
    no host synchronisation), so the whole exchange is captured inside the phase's CUDA graph."""

    def __init__(self, flat, world, bucket_mb=25.0):
        self.flat, self.world = flat, world
        self.buckets = []   # [lo, hi, [(name, param)]]
        cur, lo = [], 0
        limit = int(bucket_mb * (1 << 20) / 4)
        for name, p in flat.params:
            cur.append((name, p))
            hi = flat.offsets[name] + (p.numel() + 3) // 4 * 4
            if hi - lo >= limit:
                self.buckets.append([lo, hi, cur])
                cur, lo = [], hi
        if cur:
            self.buckets.append([lo, flat.numel, cur])
        self._of = {}
        for bi, (_, _, ps) in enumerate(self.buckets):
            for _, p in ps:
                self._of[id(p)] = bi
        self._pending = [0] * len(self.buckets)
        self._flushed = [True] * len(self.buckets)
        self._arrived = [[] for _ in self.buckets]
        self._held = []
        self._active = False
        self._side = None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for _, p in flat.params]

    def begin(self):
        """Call right before backward(): zero the flat buffer, drop stale .grad tensors, arm the buckets."""
        self.flat.grad.zero_()
        for _, p in self.flat.params:
            p.grad = None
        for bi, (_, _, ps) in enumerate(self.buckets):
            self._pending[bi] = sum(1 for _, p in ps if p.requires_grad)
            self._flushed[bi] = False
            self._arrived[bi] = []
        self._held = []
        self._active = True
        if self.flat.grad.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(self.flat.grad.device)
            # the zeroing above precedes every bucket's copy on the side stream
            self._side.wait_stream(torch.cuda.current_stream(self.flat.grad.device))

    def _on_grad(self, p):
        if not self._active:
            return
        bi = self._of[id(p)]
        if p.grad is not None and p.grad.is_cuda:
            # autograd runs AccumulateGrad on the stream the parameter was first used on (the generator's style work
            # lives on a side stream): remember where this gradient became final
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(p.grad.device))
            self._arrived[bi].append(ev)
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and not self._flushed[bi]:
            self._flush(bi)

    def _flush(self, bi):
        lo, hi, ps = self.buckets[bi]
        self._flushed[bi] = True
        views, grads = [], []
        for name, p in ps:
            if p.grad is not None:
                o = self.flat.offsets[name]
                views.append(self.flat.grad[o:o + p.numel()].view(p.shape))
                grads.append(p.grad)
            p.grad = None
        chunk = self.flat.grad[lo:hi]
        if chunk.is_cuda:
            # copy + all-reduce on the side stream, after every gradient of the bucket is final; the gradient tensors
            # stay referenced until finish() so that their memory is not handed out again while the side stream reads it
            self._held.append(grads)
            for ev in self._arrived[bi]:
                self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                if grads:
                    torch._foreach_copy_(views, grads)
                if self.world > 1:
                    dist.all_reduce(chunk, op=dist.ReduceOp.SUM)
        else:
            if grads:
                torch._foreach_copy_(views, grads)
            if self.world > 1:
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM)

    def finish(self):
        """Call after backward(): flush the buckets that never filled, wait for the exchanges."""
        for bi in range(len(self.buckets) - 1, -1, -1):
            if not self._flushed[bi]:
                self._flush(bi)
        self._active = False
        if self._side is not None:
            torch.cuda.current_stream(self.flat.grad.device).wait_stream(self._side)
        self._held = []


class FlatAdam:
    """torch.optim.Adam semantics on a FlatParams through te_adam_ema."""

    def __init__(self, flat, lr, betas, eps=1e-8):
        self.flat = flat
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = torch.zeros_like(flat.data)
        self.v = torch.zeros_like(flat.data)
        # per-range step counters (Adam's bias correction counts the updates a parameter received);
        # they live on the DEVICE so that the update can be replayed from a captured CUDA graph
        self.steps = [torch.zeros(1, dtype=torch.int32, device=flat.data.device)
                      for _ in flat.group_end]

    def step(self, n_groups, grad_scale=1.0, ema=None, ema_decay=0.0):
        """Update parameter groups [0, n_groups) (the main group is group 0).  `ema` (a flat buffer laid out like
        the parameters) receives ema = ema_decay * ema + (1 - ema_decay) * p_new in the same launch: the reference's
        `accumulate` loop (train_spatial_query.py:56-61) without a pass of its own.  Groups that are not stepped
        keep p == const, so their EMA entries are left alone."""
        lo = 0
        for gi in range(n_groups):
            hi = self.flat.group_end[gi]
            if hi > lo:
                self.steps[gi].add_(1)
                sl = slice(lo, hi)
                lib.adam_ema_devstep(self.flat.data[sl], self.flat.grad[sl], self.m[sl], self.v[sl],
                                     None if ema is None else ema[sl],
                                     self.lr, self.betas[0], self.betas[1], self.eps, self.steps[gi], ema_decay,
                                     grad_scale)
            lo = hi


def d_logistic_loss(real_pred, fake_pred):
    """train_spatial_query.py:69-73"""
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    """train_spatial_query.py:76-83"""
    (grad_real,) = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_nonsaturating_loss(fake_pred):
    """train_spatial_query.py:86-89"""
    return F.softplus(-fake_pred).mean()


def g_path_regularize(fake_img, latents, mean_path_length, decay=0.01):
    """train_spatial_query.py:92-105"""
    noise = torch.randn_like(fake_img) / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    (grad,) = torch.autograd.grad(outputs=(fake_img * noise).sum(), inputs=latents, create_graph=True)
    path_lengths = torch.sqrt(grad.pow(2).sum(2).mean(1))
    path_mean = mean_path_length + decay * (path_lengths.mean() - mean_path_length)
    path_penalty = (path_lengths - path_mean).pow(2).mean()
    return path_penalty, path_mean.detach(), path_lengths


def _set_requires_grad(flat, flag):
    for _, p in flat.params:
        p.requires_grad_(flag)


class Trainer:
    def __init__(self, cfg, device, seed=0):
        self.cfg = cfg
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        torch.manual_seed(seed)  # identical initial weights on every rank
        mk = lambda: Generator(cfg.size, cfg.latent, cfg.latent, cfg.token,  # noqa: E731
                               channel_multiplier=cfg.channel_multiplier,
                               layer_noise_injection=cfg.inject_noise, n_trans=cfg.num_trans,
                               pixel_norm_op_dim=cfg.pixel_norm_op_dim)
        self.generator = mk().to(self.device)
        self.discriminator = Discriminator(cfg.size, channel_multiplier=cfg.channel_multiplier).to(self.device)
        self.g_ema = mk().to(self.device).eval()
        rgb_bias = lambda n: n.startswith("to_rgb") and n.endswith(".bias") and ".conv." not in n  # noqa: E731
        noise_w = lambda n: n.endswith("noise.weight")  # noqa: E731
        tails = [rgb_bias] if cfg.inject_noise else [rgb_bias, noise_w]
        self.g_flat = FlatParams(self.generator, tails)
        self.d_flat = FlatParams(self.discriminator)
        self.ema_flat = FlatParams(self.g_ema, tails)
        self.ema_flat.data.copy_(self.g_flat.data)  # accumulate(g_ema, generator, 0), :456
        for _, p in self.ema_flat.params:
            p.requires_grad_(False)
            p.grad = None
        self.ema_flat.grad = None
        g_ratio = cfg.g_reg_every / (cfg.g_reg_every + 1)
        d_ratio = cfg.d_reg_every / (cfg.d_reg_every + 1)
        self.g_optim = FlatAdam(self.g_flat, cfg.lr * g_ratio, (0 ** g_ratio, 0.99 ** g_ratio))
        self.d_optim = FlatAdam(self.d_flat, cfg.lr * d_ratio, (0 ** d_ratio, 0.99 ** d_ratio))
        torch.manual_seed(1234 + self.rank)  # per-rank latent / noise streams
        self.mean_path_length = torch.zeros((), device=self.device)
        self.mean_spatial_path_length = torch.zeros((), device=self.device)
        self.iteration = 0
        self.losses = {}
        self.use_graphs = False
        self._graphs = {}
        self._real = None  # static input buffer (graph replays read from it)
        # bf16 route: tap-major bf16 copies of the shared conv weights, re-packed once per optimiser step
        # (consulted only inside this trainer's own phases: tc.use_pack_cache)
        self._packs = tc.PackCache()
        tc.clear_wgrad_workspaces()
        self._packs.register(p for _, p in self.g_flat.params + self.d_flat.params)
        self._ema_in_step = False  # set per phase by step(): the LAST generator update of an iteration carries the EMA
        # gradient exchange overlapped with backward (N > 1): bucketed all-reduces issued from autograd hooks
        # (TE_GRAD_BUCKETS=0: one flat all-reduce after the backward pass instead; TE_GRAD_BUCKET_MB sizes the buckets)
        overlap = self.world > 1 and os.environ.get("TE_GRAD_BUCKETS", "1") != "0"
        mb = float(os.environ.get("TE_GRAD_BUCKET_MB", "25"))
        self.g_buckets = GradBuckets(self.g_flat, self.world, mb) if overlap else None
        self.d_buckets = GradBuckets(self.d_flat, self.world, mb) if overlap else None

    def weights_changed(self):
        """Call after modifying generator / discriminator weights outside this trainer's optimiser steps
        (load_state_dict, manual edits): re-packs the cached bf16 weight copies, which replayed CUDA graphs read
        without consulting version counters."""
        self._packs.refresh()

    def _step_optim(self, flat, optim, n_groups, with_ema=False):
        if with_ema:
            optim.step(n_groups, grad_scale=1.0 / self.world, ema=self.ema_flat.data, ema_decay=self.cfg.ema_decay)
        else:
            optim.step(n_groups, grad_scale=1.0 / self.world)
        lo = flat.data.data_ptr()
        self._packs.refresh(lo, lo + flat.data.numel() * flat.data.element_size())

    # ------------------------------------------------------------------ CUDA graphs
    def enable_graphs(self, flag=True):
        """Replay each phase (D step, R1, G step, path regulariser) from captured CUDA graphs: forward + backward
        (incl. the bucketed NCCL gradient exchange when N > 1) in one graph, the optimiser update in a second one.
        Call after at least one eager iteration (lazy initialisation must be done)."""
        self.use_graphs = flag
        if not flag:
            self._graphs = {}   # captured graphs (and the buffers in their pools) go; a later enable re-captures
        tc.clear_wgrad_workspaces()

    def _phase(self, name, fwdbwd, flat, optim, n_groups):
        with tc.use_pack_cache(self._packs):
            self._phase_inner(name, fwdbwd, flat, optim, n_groups)

    def _phase_inner(self, name, fwdbwd, flat, optim, n_groups):
        with_ema = self._ema_in_step and flat is self.g_flat
        if not self.use_graphs:
            fwdbwd()
            self._reduce_and_step(flat, optim, n_groups, with_ema)
            return
        entry = self._graphs.get(name)
        if entry is None:
            torch.cuda.synchronize()
            n0 = lib.launch_count
            ga = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                fwdbwd()
            entry = self._graphs[name] = [ga, {}, lib.launch_count - n0]
        ga, steps, launches = entry
        if with_ema not in steps:  # the optimiser graph exists in two flavours: with and without the fused EMA
            torch.cuda.synchronize()
            n0 = lib.launch_count
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb):
                self._step_optim(flat, optim, n_groups, with_ema)
            steps[with_ema] = (gb, lib.launch_count - n0)
            # capturing does not execute: undo the step counters' host-side bookkeeping is not needed (device
            # counters are only incremented by the captured add_ when the graph runs)
        gb, step_launches = steps[with_ema]
        ga.replay()   # N > 1: the bucketed NCCL all-reduces are part of the captured graph
        gb.replay()
        lib.launch_count += launches + step_launches

    # ------------------------------------------------------------------ pieces of one iteration
    def _latents(self, n):
        c = self.cfg
        z = torch.randn(n, c.latent, c.para_num, device=self.device)  # utils/sample.py:19
        p = torch.randn(n, c.latent, c.para_num, device=self.device)  # utils/sample.py:10
        return z, p

    def _backward(self, loss, flat, buckets):
        """loss.backward() with the gradients landing in `flat.grad` — summed over ranks when N > 1 (bucketed
        all-reduces overlapped with the backward pass, GradBuckets); N = 1: one flat gather at the end."""
        if buckets is None:
            flat.clear_grads()
            loss.backward()
            flat.gather_grads()
            if self.world > 1:
                dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
            return
        buckets.begin()
        loss.backward()
        buckets.finish()

    def _reduce_and_step(self, flat, optim, n_groups, with_ema=False):
        self._step_optim(flat, optim, n_groups, with_ema)

    def _d_fwdbwd(self):
        _set_requires_grad(self.g_flat, False)
        _set_requires_grad(self.d_flat, True)
        z, p = self._latents(self.cfg.batch)
        fake_img, _, _ = self.generator(z, p)
        # one discriminator pass over [fake; real] (two sub-batches): same logits as two calls
        # (train_spatial_query.py:199-201), half the launches and weight repacks
        fake_pred, real_pred = self.discriminator.forward_stacked(torch.cat([fake_img, self._real]), 2).chunk(2)
        d_loss = d_logistic_loss(real_pred, fake_pred)
        self._backward(d_loss, self.d_flat, self.d_buckets)
        self.losses.update(d=d_loss.detach(), real_score=real_pred.mean().detach(),
                           fake_score=fake_pred.mean().detach())

    def _dreg_fwdbwd(self):
        _set_requires_grad(self.g_flat, False)
        _set_requires_grad(self.d_flat, True)
        real_img = self._real.detach().requires_grad_(True)
        real_pred = self.discriminator(real_img)
        r1_loss = d_r1_loss(real_pred, real_img)
        self._backward(self.cfg.r1 / 2 * r1_loss * self.cfg.d_reg_every + 0 * real_pred[0], self.d_flat, self.d_buckets)
        self.losses["r1"] = r1_loss.detach()

    def _g_fwdbwd(self):
        _set_requires_grad(self.g_flat, True)
        _set_requires_grad(self.d_flat, False)
        z, p = self._latents(self.cfg.batch)
        fake_img, _, _ = self.generator(z, p)
        g_loss = g_nonsaturating_loss(self.discriminator(fake_img))
        self._backward(g_loss, self.g_flat, self.g_buckets)
        self.losses["g"] = g_loss.detach()

    def _greg_fwdbwd(self):
        c = self.cfg
        _set_requires_grad(self.g_flat, True)
        _set_requires_grad(self.d_flat, False)
        n = max(1, c.batch // c.path_batch_shrink)
        z, p = self._latents(n)
        fake_img, latents, _ = self.generator(z, p, return_latents=True)
        path_loss, path_mean, path_lengths = g_path_regularize(fake_img, latents, self.mean_path_length)
        weighted = c.path_regularize * c.g_reg_every * path_loss
        if c.path_batch_shrink:
            weighted = weighted + 0 * fake_img[0, 0, 0, 0]
        self._backward(weighted, self.g_flat, self.g_buckets)
        self.mean_path_length.copy_(path_mean)  # in place: a static buffer for graph replays
        self.losses.update(path=path_loss.detach(), path_length=path_lengths.mean().detach())

    def _gsreg_fwdbwd(self):
        """Spatial path regulariser (train_spatial_query.py:252-277): the path-length penalty taken with respect to
        the spatial code — the raw p ("p") or the mapped p+ ("p+") — instead of the style latents.  Second order
        runs through the const-input convolutions, the cross-attention stack (as queries) and, for "p", the spatial
        mapping network."""
        c = self.cfg
        _set_requires_grad(self.g_flat, True)
        _set_requires_grad(self.d_flat, False)
        n = max(1, c.batch // c.path_batch_shrink)
        z, p = self._latents(n)
        if c.regu_space == "p":
            target = p.requires_grad_()
            fake_img, _, _ = self.generator(z, target)
        else:
            target = self.generator(z, p, return_only_mapped_p=True)
            target.requires_grad_()
            fake_img, _, _ = self.generator(z, target, use_spatial_mapping=False)
        path_loss, path_mean, path_lengths = g_path_regularize(fake_img, target, self.mean_spatial_path_length)
        weighted = c.spatial_path_regularize * c.g_reg_every * path_loss
        if c.path_batch_shrink:
            weighted = weighted + 0 * fake_img[0, 0, 0, 0]
        self._backward(weighted, self.g_flat, self.g_buckets)
        self.mean_spatial_path_length.copy_(path_mean)
        self.losses.update(spatial_path=path_loss.detach(), spatial_path_length=path_lengths.mean().detach())

    def _set_real(self, real_img):
        if real_img is self._real:
            return
        if self._real is None or self._real.shape != real_img.shape:
            if self._graphs:
                # captured graphs read the OLD static buffer at the OLD batch size: refuse instead of silently
                # training on stale data (a partial last batch must be dropped or padded by the caller)
                raise RuntimeError("Trainer: the image batch changed shape from %s to %s after CUDA graphs were "
                                   "captured; call enable_graphs(False) or keep the batch shape fixed"
                                   % (tuple(self._real.shape), tuple(real_img.shape)))
            self._real = torch.empty(real_img.shape, dtype=torch.float32, device=self.device)
        self._real.copy_(real_img, non_blocking=True)

    def d_step(self, real_img):
        self._set_real(real_img)
        self._phase("d", self._d_fwdbwd, self.d_flat, self.d_optim, 1)

    def d_regularize(self, real_img):
        self._set_real(real_img)
        self._phase("dreg", self._dreg_fwdbwd, self.d_flat, self.d_optim, 1)

    def g_step(self):
        # noise strengths receive no gradient when noise injection is off (reference: grad None)
        self._phase("g", self._g_fwdbwd, self.g_flat, self.g_optim, 2)

    def g_regularize(self):
        # The reference adds `0 * fake_img[0,0,0,0]` to the path penalty (train_spatial_query.py:240-243), so the
        # to_rgb biases receive a ZERO gradient tensor (not None) and torch's Adam steps them: step += 1 and
        # exp_avg_sq *= beta2 with no parameter change (beta1 = 0).  The gathered flat gradient is zero there, so
        # stepping group 1 reproduces exactly that; only the noise strengths (grad None) stay skipped.
        self._phase("greg", self._greg_fwdbwd, self.g_flat, self.g_optim, 2)

    def g_spatial_regularize(self):
        self._phase("gsreg", self._gsreg_fwdbwd, self.g_flat, self.g_optim, 2)

    def ema_update(self):
        """accumulate(g_ema, g_module, 0.5 ** (32 / 10000)), train_spatial_query.py:56-61,294 as a stand-alone
        pass (step() fuses it into the last generator update instead)."""
        self.ema_flat.data.lerp_(self.g_flat.data, 1.0 - self.cfg.ema_decay)

    # ------------------------------------------------------------------ one iteration
    def step(self, real_img):
        """real_img: [B,3,S,S] on the device, range [-1,1]."""
        i = self.iteration
        self.d_step(real_img)
        if i % self.cfg.d_reg_every == 0:
            self.d_regularize(real_img)
        # the EMA (accumulate(g_ema, g, decay) after the iteration's generator updates, :294) rides in the optimiser
        # kernel of the LAST generator phase of this iteration
        greg = i % self.cfg.g_reg_every == 0
        try:
            self._ema_in_step = not greg
            self.g_step()
            if greg:
                self._ema_in_step = not self.cfg.spatial_regu
                self.g_regularize()
                if self.cfg.spatial_regu:
                    self._ema_in_step = True
                    self.g_spatial_regularize()
        finally:
            self._ema_in_step = False
        self.iteration += 1
        return self.losses

    def reduced_losses(self):
        """The loss dictionary averaged over ranks on rank 0 — `reduce_loss_dict` (utils/distributed.py:102-124, called
        at train_spatial_query.py:296): one stacked [n] vector, dist.reduce(dst=0), divided by the world size there;
        the other ranks keep their local values, like the reference.  Returns (keys, stacked device tensor)."""
        keys = sorted(self.losses)
        vec = torch.stack([self.losses[k].float() for k in keys])
        if self.world > 1:
            dist.reduce(vec, dst=0)
            if self.rank == 0:
                vec = vec / self.world
        return keys, vec

    def mean_path_length_avg(self):
        """`reduce_sum(mean_path_length).item() / world_size` (train_spatial_query.py:248-250): the logged average of
        the ranks' running path-length means (each rank keeps its own running mean, as in the reference)."""
        v = self.mean_path_length.detach().clone()
        if self.world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return v / self.world

    def step_from_host_uint8(self, u8_pinned, flip_pinned=None):
        """End-to-end form on the data path's own format: decoded uint8 [B, H, W, 3] pixels (pinned) and the
        RandomHorizontalFlip coins in, host loss scalars out.  The H2D copy carries a quarter of the float32 bytes;
        mirror / ToTensor / Normalize (train_spatial_query.py:511-517) run on the device in one kernel that writes
        straight into the static input buffer the captured graphs read."""
        from .data import DeviceImagePipeline
        b, h, w, _ = u8_pinned.shape
        if getattr(self, "_pipe", None) is None or self._pipe.u8.shape != u8_pinned.shape:
            self._pipe = DeviceImagePipeline(b, h, self.device)
        if self._real is None or self._real.shape != (b, 3, h, w):
            self._set_real(torch.empty((b, 3, h, w), dtype=torch.float32, device=self.device))
        self._pipe.load(u8_pinned, flip_pinned, out=self._real)
        self.step(self._real)
        keys, vec = self.reduced_losses()
        return dict(zip(keys, vec.cpu().tolist()))

    def step_from_host(self, real_pinned):
        """End-to-end form: host (pinned) images in, host loss scalars out (rank-averaged on rank 0)."""
        self._set_real(real_pinned)  # H2D straight into the static input buffer
        self.step(self._real)
        keys, vec = self.reduced_losses()
        host = vec.cpu()  # D2H + sync, like the .item()s at :298-306
        return dict(zip(keys, host.tolist()))
