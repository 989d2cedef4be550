"""Dual-space inversion forward (BASELINE configs[4]): image -> (Z+, P+) codes through the pSp
`GradualStyleEncoder` (IR-SE50 trunk + 14 style heads + 16 spatial heads), then the generator with both mapping
networks bypassed.

Reference interfaces mirrored (same constructor arguments, parameter names and state_dict keys, so the authors' pSp
checkpoints — keys `encoder.*` / `decoder.*` — load unchanged):
  pSp/models/encoders/helpers.py:25-120        get_blocks, SEModule, bottleneck_IR, bottleneck_IR_SE
  pSp/models/encoders/psp_encoders_new.py:13-140  GradualStyleBlock, GradualStyleEncoder
  pSp/models/psp_new.py:90-131                 pSp.forward / only_decode (encode, add the average code, decode, pool)
  dual_space_encoder.py:12-33                  DualSpaceEncoder.encode / decode

SURVEY.md §8d lists the encoder of config 5 as stock PyTorch; the module classes here are exactly that (any device,
any dtype).  `FusedEncoder` is the B200 inference form:
  * the 30 heads (16 spatial, 3/4/7 style per pyramid level) are 142 stride-2 3x3 convolutions 512 -> 512 — 44 % of
    the encoder's FLOPs: in bf16 they run on this package's tcgen05 convolution kernel (bias + LeakyReLU in the
    epilogue), and each group's EqualLinears are one batched product;
  * the IR-SE50 trunk's 3x3 and strided 1x1 convolutions run on the same kernel, per-channel PReLU and the folded
    batch-norm shift in its epilogue (eval-mode batch norms that FOLLOW a convolution are folded into it; the one in
    front of conv1 stays a fused multiply-add because zero padding comes after it); bf16 channels-last activations;
    stem, lateral 1x1 layers and squeeze-excite gates are library / elementwise calls;
  * encoder + generator are captured into one CUDA graph for fixed shapes (`InversionPipeline`).
The generator half runs on this package's own kernels (bf16 tcgen05 route).
"""
import math
from collections import namedtuple

import torch
from torch import nn
from torch.nn import functional as F

from .model import EqualLinear

Bottleneck = namedtuple("Bottleneck", ["in_channel", "depth", "stride"])


def get_blocks(num_layers):
    """Unit list of the IR trunk (helpers.py:29-57): (units per stage) x (in, depth), first unit of a stage strided."""
    stages = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}
    if num_layers not in stages:
        raise ValueError("Invalid number of layers: {}. Must be one of [50, 100, 152]".format(num_layers))
    blocks, in_channel = [], 64
    for units, depth in zip(stages[num_layers], (64, 128, 256, 512)):
        blocks.append([Bottleneck(in_channel, depth, 2)] + [Bottleneck(depth, depth, 1) for _ in range(units - 1)])
        in_channel = depth
    return blocks


class SEModule(nn.Module):
    """helpers.py:60-77 — squeeze (global average) and excite (two 1x1 convolutions, sigmoid gate)."""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        gate = self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))
        return x * gate


class bottleneck_IR(nn.Module):
    """helpers.py:80-98"""
    _with_se = False

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False),
                                                nn.BatchNorm2d(depth))
        layers = [nn.BatchNorm2d(in_channel), nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False),
                  nn.PReLU(depth), nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False), nn.BatchNorm2d(depth)]
        if self._with_se:
            layers.append(SEModule(depth, 16))
        self.res_layer = nn.Sequential(*layers)

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class bottleneck_IR_SE(bottleneck_IR):
    """helpers.py:101-120"""
    _with_se = True


class GradualStyleBlock(nn.Module):
    """psp_encoders_new.py:13-32 — log2(spatial) stride-2 convolutions down to 1x1, then an EqualLinear."""

    def __init__(self, in_c, out_c, spatial):
        super().__init__()
        self.out_c = out_c
        self.spatial = spatial
        modules = []
        for i in range(int(math.log2(spatial))):
            modules += [nn.Conv2d(in_c if i == 0 else out_c, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
        self.convs = nn.Sequential(*modules)
        self.linear = EqualLinear(out_c, out_c, lr_mul=1)

    def forward(self, x):
        return self.linear(self.convs(x).view(-1, self.out_c))


class GradualStyleEncoder(nn.Module):
    """psp_encoders_new.py:35-140.  forward(x [B, input_nc, 256, 256]) -> (z [B, 512, 16], p [B, 512, 16])."""

    def __init__(self, num_layers, mode="ir", opts=None):
        super().__init__()
        assert num_layers in [50, 100, 152], "num_layers should be 50,100, or 152"
        assert mode in ["ir", "ir_se"], "mode should be ir or ir_se"
        unit = bottleneck_IR if mode == "ir" else bottleneck_IR_SE
        input_nc = getattr(opts, "input_nc", 3) if opts is not None else 3
        self.input_layer = nn.Sequential(nn.Conv2d(input_nc, 64, (3, 3), 1, 1, bias=False), nn.BatchNorm2d(64),
                                         nn.PReLU(64))
        self.body = nn.Sequential(*[unit(b.in_channel, b.depth, b.stride) for blk in get_blocks(num_layers) for b in blk])
        self.styles = nn.ModuleList()
        self.style_count = 14
        self.coarse_ind = 3
        self.middle_ind = 7
        self.spatial_count = 16
        self.spatials = nn.ModuleList()
        for i in range(self.style_count):
            self.styles.append(GradualStyleBlock(512, 512, 16 if i < self.coarse_ind else 32 if i < self.middle_ind else 64))
        for _ in range(self.spatial_count):
            self.spatials.append(GradualStyleBlock(512, 512, 16))
        self.latlayer1 = nn.Conv2d(256, 512, kernel_size=1, stride=1, padding=0)
        self.latlayer2 = nn.Conv2d(128, 512, kernel_size=1, stride=1, padding=0)
        self.adjust_style = EqualLinear(in_dim=14, out_dim=16)

    @staticmethod
    def _upsample_add(x, y):
        return F.interpolate(x, size=y.shape[2:], mode="bilinear", align_corners=True) + y

    def pyramid(self, x):
        """The three feature maps the heads read: c3 (512 ch, /16), p2 (512 ch, /8), p1 (512 ch, /4) (:101-124)."""
        x = self.input_layer(x)
        taps = {}
        for i, unit in enumerate(self.body):
            x = unit(x)
            if i in (6, 20, 23):
                taps[i] = x
        c1, c2, c3 = taps[6], taps[20], taps[23]
        p2 = self._upsample_add(c3, self.latlayer1(c2))
        p1 = self._upsample_add(p2, self.latlayer2(c1))
        return c3, p2, p1

    def forward(self, x):
        c3, p2, p1 = self.pyramid(x)
        feats = [c3] * self.coarse_ind + [p2] * (self.middle_ind - self.coarse_ind) + \
                [p1] * (self.style_count - self.middle_ind)
        z = torch.stack([head(f) for head, f in zip(self.styles, feats)], dim=1)        # [B, 14, 512]
        z_out = self.adjust_style(z.permute(0, 2, 1))                                  # [B, 512, 16]
        p_out = torch.stack([head(c3) for head in self.spatials], dim=1).permute(0, 2, 1)  # [B, 512, 16]
        return z_out, p_out


# ------------------------------------------------------------------------------------------------ inference form
class _HeadGroup:
    """Identical GradualStyleBlocks that read the same feature map, run together.

    bf16 on a CUDA device: every 3x3 stride-2 convolution goes through this package's tcgen05 kernel
    (`tc.conv_raw`, geometry `down1`, bias + LeakyReLU(0.01) in the epilogue), weights packed once.  (cuDNN's
    grouped-convolution kernel, the obvious way to batch the heads, runs at ~2 % of the tensor peak here:
    88 of the encoder's 100 ms at batch 32.)  Other dtypes / devices: layer 0 as ONE convolution with the heads'
    output channels concatenated, later layers as grouped convolutions.  The EqualLinears: one batched product."""

    def __init__(self, heads, dtype, channels_last=True):
        n, c = len(heads), heads[0].out_c
        convs = [[m for m in h.convs if isinstance(m, nn.Conv2d)] for h in heads]
        self.n, self.c = n, c
        self.slope = heads[0].convs[1].negative_slope
        dev = convs[0][0].weight.device
        self.use_tc = dtype == torch.bfloat16 and dev.type == "cuda" and abs(self.slope - 0.01) < 1e-12
        self.weights, self.biases = [], []
        if self.use_tc:
            from . import tc
            self._mode = tc.Mode("down1", 3)
            for cv in convs:   # per head: [(packed bf16 [9, 512, 512], f32 bias)] per layer
                self.weights.append([tc.pack_weight(m.weight.detach(), False) for m in cv])
                self.biases.append([m.bias.detach().float().contiguous() for m in cv])
        else:
            for layer in range(len(convs[0])):
                w = torch.cat([cv[layer].weight.detach() for cv in convs], 0).to(dtype)
                if channels_last:
                    w = w.contiguous(memory_format=torch.channels_last)
                self.weights.append(w)
                self.biases.append(torch.cat([cv[layer].bias.detach() for cv in convs], 0).to(dtype))
        lin = [h.linear for h in heads]
        self.lin_w = torch.stack([l.weight.detach() * l.scale for l in lin]).to(dtype)        # [H, out, in]
        self.lin_b = torch.stack([l.bias.detach() * l.lr_mul for l in lin]).to(dtype)          # [H, out]

    def __call__(self, feat):
        if self.use_tc:
            from . import tc
            outs = []
            for ws, bs in zip(self.weights, self.biases):
                x = feat
                for w, b in zip(ws, bs):
                    x = tc.conv_raw(x, w, self._mode, bias=b, act=2)
                outs.append(x.reshape(x.shape[0], self.c))
            x = torch.stack(outs)                                                     # [H, B, 512]
        else:
            x = feat
            for layer, (w, b) in enumerate(zip(self.weights, self.biases)):
                x = F.leaky_relu(F.conv2d(x, w, b, stride=2, padding=1, groups=1 if layer == 0 else self.n), self.slope)
            x = x.reshape(x.shape[0], self.n, self.c).transpose(0, 1)                 # [H, B, 512]
        y = torch.baddbmm(self.lin_b.unsqueeze(1), x, self.lin_w.transpose(1, 2))     # [H, B, 512]
        return y.transpose(0, 1)                                                      # [B, H, 512]


def _fold_bn(conv_w, bn):
    """(w', b') with  bn(conv(x, w)) == conv(x, w') + b'  for an eval-mode batch norm after a bias-free convolution."""
    inv = torch.rsqrt(bn.running_var + bn.eps) * bn.weight
    return conv_w * inv.view(-1, 1, 1, 1), bn.bias - bn.running_mean * inv


class FusedEncoder:
    """Inference form of a GradualStyleEncoder in eval mode: folded batch norms, batched heads, `dtype`
    channels-last activations.  Snapshot of the weights at construction (rebuild after loading new ones)."""

    def __init__(self, enc, dtype=torch.bfloat16):
        if enc.training:
            raise RuntimeError("FusedEncoder needs the encoder in eval mode (batch norms are folded)")
        self.dtype = dtype
        dev = enc.input_layer[0].weight.device
        # bf16 on a CUDA device: the trunk's 3x3 / strided 1x1 convolutions run on the tcgen05 kernel too (PReLU and
        # the folded batch-norm shift in its epilogue); the 3-channel stem, the two lateral 1x1 layers and the
        # squeeze-excite arithmetic stay library / elementwise calls
        self.use_tc = dtype == torch.bfloat16 and dev.type == "cuda"
        cl = lambda w: w.to(dtype).contiguous(memory_format=torch.channels_last)  # noqa: E731
        if self.use_tc:
            from . import tc
            conv_w = lambda w: tc.pack_weight(w.detach(), False)  # noqa: E731
        else:
            conv_w = lambda w: cl(w.detach())  # noqa: E731
        with torch.no_grad():
            w, b = _fold_bn(enc.input_layer[0].weight, enc.input_layer[1])
            self.stem = (cl(w), b.to(dtype), enc.input_layer[2].weight.detach().to(dtype))
            self.units = []
            for u in enc.body:
                r = u.res_layer
                bn1 = r[0]
                inv1 = torch.rsqrt(bn1.running_var + bn1.eps) * bn1.weight
                w2, b2 = _fold_bn(r[3].weight, r[4])
                unit = {"bn1_scale": inv1.to(dtype).view(1, -1, 1, 1),
                        "bn1_shift": (bn1.bias - bn1.running_mean * inv1).to(dtype).view(1, -1, 1, 1),
                        "w1": conv_w(r[1].weight),
                        "prelu": r[2].weight.detach().float().contiguous() if self.use_tc else r[2].weight.detach().to(dtype),
                        "w2": conv_w(w2), "b2": b2.float().contiguous() if self.use_tc else b2.to(dtype),
                        "stride": r[3].stride[0], "se": None, "short": None}
                if len(r) > 5:
                    unit["se"] = (cl(r[5].fc1.weight.detach()), cl(r[5].fc2.weight.detach()))
                if isinstance(u.shortcut_layer, nn.Sequential):
                    ws, bs = _fold_bn(u.shortcut_layer[0].weight, u.shortcut_layer[1])
                    unit["short"] = (conv_w(ws), bs.float().contiguous() if self.use_tc else bs.to(dtype))
                self.units.append(unit)
            self.lat1 = (cl(enc.latlayer1.weight.detach()), enc.latlayer1.bias.detach().to(dtype))
            self.lat2 = (cl(enc.latlayer2.weight.detach()), enc.latlayer2.bias.detach().to(dtype))
            s = list(enc.styles)
            self.coarse = _HeadGroup(s[:enc.coarse_ind], dtype)
            self.middle = _HeadGroup(s[enc.coarse_ind:enc.middle_ind], dtype)
            self.fine = _HeadGroup(s[enc.middle_ind:], dtype)
            self.spatial = _HeadGroup(list(enc.spatials), dtype)
            a = enc.adjust_style
            self.adjust = ((a.weight.detach() * a.scale).to(dtype), (a.bias.detach() * a.lr_mul).to(dtype))

    def _unit(self, x, u):
        s = u["stride"]
        # the batch norm in FRONT of conv1 cannot be folded (zero padding follows it): one fused multiply-add
        r = torch.addcmul(u["bn1_shift"], x, u["bn1_scale"])
        if self.use_tc:
            from . import tc
            if u["short"] is not None:
                short = tc.conv_raw(x, u["short"][0], tc.Mode("down" if s == 2 else "s1", 1), bias=u["short"][1])
            else:
                short = x if s == 1 else x[:, :, ::s, ::s]
            r = tc.conv_raw(r, u["w1"], tc.Mode("s1", 3), act=3, slope=u["prelu"])
            r = tc.conv_raw(r, u["w2"], tc.Mode("down1" if s == 2 else "s1", 3), bias=u["b2"])
        else:
            if u["short"] is not None:
                short = F.conv2d(x, u["short"][0], u["short"][1], stride=s)
            else:
                short = x if s == 1 else x[:, :, ::s, ::s]      # MaxPool2d(1, stride) == strided sampling
            r = F.prelu(F.conv2d(r, u["w1"], None, stride=1, padding=1), u["prelu"])
            r = F.conv2d(r, u["w2"], u["b2"], stride=s, padding=1)
        if u["se"] is not None:
            g = torch.sigmoid(F.conv2d(F.relu(F.conv2d(r.mean((2, 3), keepdim=True), u["se"][0])), u["se"][1]))
            return torch.addcmul(short, r, g)
        return r + short

    @torch.no_grad()
    def __call__(self, x):
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = F.prelu(F.conv2d(x, self.stem[0], self.stem[1], padding=1), self.stem[2])
        taps = {}
        for i, u in enumerate(self.units):
            x = self._unit(x, u)
            if i in (6, 20, 23):
                taps[i] = x
        c1, c2, c3 = taps[6], taps[20], taps[23]
        up = lambda a, b: F.interpolate(a, size=b.shape[2:], mode="bilinear", align_corners=True) + b  # noqa: E731
        p2 = up(c3, F.conv2d(c2, *self.lat1))
        p1 = up(p2, F.conv2d(c1, *self.lat2))
        z = torch.cat([self.coarse(c3), self.middle(p2), self.fine(p1)], 1)            # [B, 14, 512]
        z_out = torch.addmm(self.adjust[1], z.transpose(1, 2).reshape(-1, 14), self.adjust[0].t())
        z_out = z_out.reshape(x.shape[0], 512, 16)
        p_out = self.spatial(c3).transpose(1, 2)                                      # [B, 512, 16]
        return z_out.float(), p_out.float()


class InversionPipeline:
    """pSp.forward for the plus-space setting (psp_new.py:90-121 with from_plus_space): codes = encoder(x) (+ the
    average codes), images = generator(z+, p+, use_style_mapping=False, use_spatial_mapping=False), optionally
    pooled to 256^2.  With `graph=True` the whole forward is captured once per input shape and replayed.

        pipe = InversionPipeline(encoder, g_ema, z_avg=..., p_avg=...)
        images, z_code, p_code = pipe(real_images)        # valid until the next call when graphed
    """

    def __init__(self, encoder, generator, z_avg=None, p_avg=None, resize=True, graph=True, dtype=torch.bfloat16):
        self.encoder = FusedEncoder(encoder.eval(), dtype)
        self.generator = generator.eval()
        self.z_avg, self.p_avg = z_avg, p_avg
        self.resize = resize
        self.use_graph = graph
        self._graph = None
        self._in = None
        self._out = None

    def encode(self, x):
        z, p = self.encoder(x)
        if self.z_avg is not None:
            z = z + self.z_avg
        if self.p_avg is not None:
            p = p + self.p_avg
        return z, p

    def decode(self, z, p):
        with torch.no_grad():
            images, _, _ = self.generator(z, p, use_spatial_mapping=False, use_style_mapping=False,
                                          return_latents=False)
        if self.resize and images.shape[-1] != 256:
            images = F.adaptive_avg_pool2d(images, (256, 256))
        return images

    def _run(self, x):
        z, p = self.encode(x)
        return self.decode(z, p), z, p

    def __call__(self, x):
        if not x.is_cuda:
            raise RuntimeError("transeditor_b200: expected a CUDA tensor (the hot path has no CPU fallback)")
        if not self.use_graph:
            return self._run(x)
        if self._graph is None or self._in.shape != x.shape:
            self._in = x.detach().clone()
            for _ in range(3):
                self._run(self._in)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._run(self._in)
            self._graph, self._out = graph, out
        self._in.copy_(x, non_blocking=True)
        self._graph.replay()
        return self._out
