"""CPU train iteration of the oracle (functional restatement of train_spatial_query.py:166-306).

TEST INFRASTRUCTURE / BASELINE ONLY: used by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference` (the reference has no runnable CPU path of its own — SURVEY.md
fact 7 — so the "reference arm" of this tier is this port on the host cores).
"""
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
