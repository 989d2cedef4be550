"""CPU train iteration of the oracle (functional restatement of train_spatial_query.py:166-306).

TEST INFRASTRUCTURE / BASELINE ONLY: used by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference` (the reference has no runnable CPU path of its own — SURVEY.md
fact 7 — so the "reference arm" of this tier is this port on the host cores).
"""
import contextlib
import math

import torch

from oracle import te_oracle as O


class CpuTrainer:
    def __init__(self, size=256, cm=2, n_trans=8, batch=1, lr=0.002, seed=0):
        self.size, self.batch, self.n_trans = size, batch, n_trans
        self.g = {k: v.clone() for k, v in O.synthetic_state(O.generator_shapes(size, cm, n_trans), seed).items()}
        self.d = {k: v.clone() for k, v in O.synthetic_state(O.discriminator_shapes(size, cm), seed).items()}
        self.g_params = [k for k in self.g if not (k.endswith("kernel") or k.startswith("noises.") or k.startswith("token"))]
        self.d_params = [k for k in self.d if not k.endswith("kernel")]
        for k in self.g_params:
            self.g[k].requires_grad_(True)
        for k in self.d_params:
            self.d[k].requires_grad_(True)
        self.g_ema = {k: self.g[k].detach().clone() for k in self.g_params}
        gr, dr = 4 / 5, 16 / 17
        self.g_opt = torch.optim.Adam([self.g[k] for k in self.g_params], lr=lr * gr, betas=(0 ** gr, 0.99 ** gr))
        self.d_opt = torch.optim.Adam([self.d[k] for k in self.d_params], lr=lr * dr, betas=(0 ** dr, 0.99 ** dr))
        self.mean_path = torch.zeros(())
        self.it = 0
        torch.manual_seed(1234)

    def _fake(self, n, with_latent=False):
        z, p = torch.randn(n, 512, 16), torch.randn(n, 512, 16)
        img, lat = O.generator_forward(self.g, z, p, self.size, self.n_trans)
        return (img, lat) if with_latent else img

    def _grads(self, on_g, on_d):
        for k in self.g_params:
            self.g[k].requires_grad_(on_g)
        for k in self.d_params:
            self.d[k].requires_grad_(on_d)

    def step(self, real):
        b = self.batch
        self._grads(False, True)
        loss = O.d_logistic_loss(O.discriminator_forward(self.d, real),
                                 O.discriminator_forward(self.d, self._fake(b)))
        self.d_opt.zero_grad()
        loss.backward()
        self.d_opt.step()
        if self.it % 16 == 0:
            r = real.clone().requires_grad_(True)
            pred = O.discriminator_forward(self.d, r)
            pen = O.d_r1_penalty(pred, r)
            self.d_opt.zero_grad()
            (10 / 2 * pen * 16 + 0 * pred[0]).backward()
            self.d_opt.step()
        self._grads(True, False)
        loss = O.g_nonsaturating_loss(O.discriminator_forward(self.d, self._fake(b)))
        self.g_opt.zero_grad()
        loss.backward()
        self.g_opt.step()
        if self.it % 4 == 0:
            img, lat = self._fake(max(1, b // 2), with_latent=True)
            noise = torch.randn_like(img) / math.sqrt(img.shape[2] * img.shape[3])
            pl = O.g_path_lengths(img, lat, noise)
            mean = self.mean_path + 0.01 * (pl.mean() - self.mean_path)
            pen = (pl - mean).pow(2).mean()
            self.mean_path = mean.detach()
            self.g_opt.zero_grad()
            (2 * 4 * pen + 0 * img[0, 0, 0, 0]).backward()
            self.g_opt.step()
        with torch.no_grad():
            for k in self.g_params:
                self.g_ema[k].mul_(0.5 ** (32 / 10000)).add_(self.g[k].detach(), alpha=1 - 0.5 ** (32 / 10000))
        self.it += 1
        return float(loss.detach())


class RefClassCpuTrainer:
    """The same iteration with the reference's OWN nn.Modules (UNMODIFIED model_spatial_query.py, imported through
    oracle/ref_shim.py: pure-torch `utils.op`, `.cuda()` neutralised) and torch.optim.Adam — the closest thing to "the
    reference on the host cores" that exists, since the reference itself is CUDA-only.  Loop body restated from
    train_spatial_query.py:166-306 (losses :64-105, optimisers :461-473, EMA :56-61)."""

    def __init__(self, size=256, cm=2, n_trans=8, batch=1, lr=0.002, seed=0, device="cpu"):
        from oracle import ref_shim
        self.shim = ref_shim
        self.device = torch.device(device)
        if self.device.type == "cuda":
            # the reference exactly as it runs on a GPU: its own CUDA extensions (oracle/ref_gpu.py), cuDNN convolutions
            from oracle import ref_gpu
            ref = ref_gpu.load_reference_model()
            self._mode = contextlib.nullcontext
        else:
            ref = ref_shim.load_reference_module()
            self._mode = ref_shim.cpu_mode
        torch.manual_seed(seed)
        t = 2 * int(math.log2(size)) - 2
        with self._mode():
            self.g = ref.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=n_trans, pixel_norm_op_dim=1,
                                   layer_noise_injection=False)
            self.g_ema = ref.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=n_trans, pixel_norm_op_dim=1,
                                       layer_noise_injection=False).eval()
            self.d = ref.Discriminator(size, channel_multiplier=cm)
        self.g, self.g_ema, self.d = self.g.to(self.device), self.g_ema.to(self.device), self.d.to(self.device)
        self.g_ema.load_state_dict(self.g.state_dict())
        gr, dr = 4 / 5, 16 / 17
        self.g_opt = torch.optim.Adam(self.g.parameters(), lr=lr * gr, betas=(0 ** gr, 0.99 ** gr))
        self.d_opt = torch.optim.Adam(self.d.parameters(), lr=lr * dr, betas=(0 ** dr, 0.99 ** dr))
        self.size, self.batch = size, batch
        self.mean_path = torch.zeros((), device=self.device)
        self.it = 0

    def _z(self, n):
        return torch.randn(n, 512, 16, device=self.device)

    @staticmethod
    def _requires_grad(model, flag):
        for p in model.parameters():
            p.requires_grad = flag

    def step(self, real):
        b = self.batch
        with self._mode():
            self._requires_grad(self.g, False)
            self._requires_grad(self.d, True)
            fake, _, _ = self.g(self._z(b), self._z(b))
            loss = O.d_logistic_loss(self.d(real), self.d(fake))
            self.d.zero_grad()
            loss.backward()
            self.d_opt.step()
            if self.it % 16 == 0:
                r = real.clone().requires_grad_(True)
                pred = self.d(r)
                pen = O.d_r1_penalty(pred, r)
                self.d.zero_grad()
                (10 / 2 * pen * 16 + 0 * pred[0]).backward()
                self.d_opt.step()
            self._requires_grad(self.g, True)
            self._requires_grad(self.d, False)
            fake, _, _ = self.g(self._z(b), self._z(b))
            loss = O.g_nonsaturating_loss(self.d(fake))
            self.g.zero_grad()
            loss.backward()
            self.g_opt.step()
            if self.it % 4 == 0:
                n = max(1, b // 2)
                img, lat, _ = self.g(self._z(n), self._z(n), return_latents=True)
                noise = torch.randn_like(img) / math.sqrt(img.shape[2] * img.shape[3])
                pl = O.g_path_lengths(img, lat, noise)
                mean = self.mean_path + 0.01 * (pl.mean() - self.mean_path)
                pen = (pl - mean).pow(2).mean()
                self.mean_path = mean.detach()
                self.g.zero_grad()
                (2 * 4 * pen + 0 * img[0, 0, 0, 0]).backward()
                self.g_opt.step()
            with torch.no_grad():
                decay = 0.5 ** (32 / 10000)
                ema = dict(self.g_ema.named_parameters())
                for k, v in self.g.named_parameters():
                    ema[k].mul_(decay).add_(v.detach(), alpha=1 - decay)
        self.it += 1
        return loss.detach() if self.device.type == "cuda" else float(loss.detach())


def make_trainer(**kw):
    """(trainer, kind): the reference's own classes when an unmodified reference tree is reachable
    ("reference"), else the functional port ("port")."""
    from oracle import ref_shim
    if ref_shim.available():
        try:
            return RefClassCpuTrainer(**kw), "reference"
        except Exception as e:  # a broken tree must not take the baseline down with it
            print("reference classes unavailable (%s); using the port" % e)
    return CpuTrainer(**kw), "port"
