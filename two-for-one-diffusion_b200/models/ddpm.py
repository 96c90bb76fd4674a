"""`GaussianDiffusion` sampling half with the reference's API (models/ddpm.py:20-99, 140-161, 195-263).

`p_sample_loop` hands whole slices of the 1000-step reverse chain to ONE launch of the fused kernel
(score forward+backward and the posterior update per step, coordinates resident on chip), instead of
~300 ATen kernels and 3 host syncs per step.  Of the training half (ddpm.py:265-337) the loss EVALUATION is provided
(q_sample / p_losses / forward under no_grad, what trainer.eval_loss runs); parameter gradients are out of scope.
"""
import warnings

import torch
from torch import nn
import torch.nn.functional as F

from dff_b200 import SCHED_KEYS, DffError
from dff_b200._native import FLAG_CENTER, FLAG_CLAMPED, FLAG_NONFINITE
from utils import assert_center_zero, center_zero, cosine_beta_schedule, extract, linear_beta_schedule


class GaussianDiffusion(nn.Module):
    def __init__(self, model, features, num_atoms, timesteps=1000, loss_type="l2", objective="pred_noise",
                 beta_schedule="cosine", p2_loss_weight_gamma=0.0, p2_loss_weight_k=1, norm_factor=1,
                 loss_weights="ones", rng="torch", chunk=100):
        super().__init__()
        self.dims, self.num_atoms, self.model = 3, num_atoms, model
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.h = features.to(self.device)
        self.objective, self.loss_type, self.norm_factor = objective, loss_type, norm_factor
        self.rng, self.chunk = rng, int(chunk)           # rng: "torch" (reference's randn stream) | "philox" (in-kernel)
        if objective != "pred_noise":
            raise DffError("only objective='pred_noise' is on the sampling path of the shipped checkpoints")
        if beta_schedule == "cosine":
            betas = cosine_beta_schedule(timesteps)
        elif beta_schedule == "linear":
            betas = linear_beta_schedule(timesteps)
        else:
            raise ValueError(f"unknown beta schedule {beta_schedule}")
        self.num_timesteps = int(betas.shape[0])
        # fp64 bookkeeping, stored as fp32 buffers under the checkpoint's names (ddpm.py:52-99)
        a = 1.0 - betas
        abar = torch.cumprod(a, 0)
        abar_prev = F.pad(abar[:-1], (1, 0), value=1.0)
        post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
        table = {
            "betas": betas, "alphas_cumprod": abar, "alphas_cumprod_prev": abar_prev,
            "sqrt_alphas_cumprod": abar.sqrt(), "sqrt_one_minus_alphas_cumprod": (1.0 - abar).sqrt(),
            "log_one_minus_alphas_cumprod": (1.0 - abar).log(), "sqrt_recip_alphas_cumprod": (1.0 / abar).sqrt(),
            "sqrt_recipm1_alphas_cumprod": (1.0 / abar - 1).sqrt(), "posterior_variance": post_var,
            "posterior_log_variance_clipped": post_var.clamp(min=1e-20).log(),
            "posterior_mean_coef1": betas * abar_prev.sqrt() / (1.0 - abar),
            "posterior_mean_coef2": (1.0 - abar_prev) * a.sqrt() / (1.0 - abar),
        }
        for k, v in table.items():
            self.register_buffer(k, v.to(torch.float32))
        # the loss weights are a checkpoint key too; values only matter for training
        if loss_weights == "ones":
            w = (p2_loss_weight_k + abar / (1 - abar)) ** -p2_loss_weight_gamma
        elif loss_weights == "score_matching":
            w = 1.0 / (1 - abar)
        elif "higheruntil_" in loss_weights:
            thr = int(loss_weights.split("_")[1])
            n = len(abar)
            w = torch.tensor([n / thr] * thr + [n / (n - thr)] * (n - thr))
        elif "lower_bound" in loss_weights:
            cl = int(loss_weights.split("_")[2])
            un = (1.0 / ((1 - abar) * (1 - betas))).clip(0, cl)
            w = un / un.sum() * len(betas)
        else:
            raise Exception(f"Wrong loss_weights: {loss_weights}")
        self.register_buffer("p2_loss_weight", w.to(torch.float32))

    # ---- per-step API (kept for callers that drive the chain themselves)
    def predict_start_from_noise(self, x_t, t, noise):
        return extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t \
            - extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise

    def q_posterior(self, x_start, x_t, t):
        mean = extract(self.posterior_mean_coef1, t, x_t.shape) * x_start \
            + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t
        return (mean, extract(self.posterior_variance, t, x_t.shape),
                extract(self.posterior_log_variance_clipped, t, x_t.shape))

    def p_mean_variance(self, x, t):
        assert_center_zero(x)
        eps = self.model(x, self.h, 1.0 * t / self.num_timesteps, alphas=None)
        x_start = center_zero(self.predict_start_from_noise(x, t=t, noise=center_zero(eps)))
        return self.q_posterior(x_start=x_start, x_t=x, t=t)

    @torch.no_grad()
    def p_sample(self, x, t):
        mean, _, logvar = self.p_mean_variance(x=x, t=t)
        noise = center_zero(torch.randn_like(x))
        keep = (t != 0).to(x.dtype).view(-1, 1, 1)
        return mean + keep * (0.5 * logvar).exp() * noise

    # ---- fused chain
    def _sched_ptrs(self):
        return [getattr(self, k) for k in SCHED_KEYS]

    @torch.no_grad()
    def p_sample_loop(self, shape):
        """x_T = centred N(0,I); T fused reverse steps; same torch RNG stream as the reference when rng='torch'
        (one randn_like per step, in step order, on the model device: ddpm.py:242, :228)."""
        device = self.betas.device
        if device.type != "cuda":
            raise DffError("sampling needs the model on a CUDA device; there is no CPU fallback")
        T = self.num_timesteps
        eng = self.model.engine(shape[0])
        sched = self._sched_ptrs()
        mol = center_zero(torch.randn(shape, device=device)).contiguous()
        eng.read_flags()
        if self.rng == "philox":
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            eng.ddpm_steps(mol, T - 1, T, T, sched, noise=None, seed=seed)
        else:
            buf = torch.empty((min(self.chunk, T),) + tuple(shape), device=device)
            done = 0
            while done < T:
                n = min(self.chunk, T - done)
                for s in range(n):
                    torch.randn(shape, device=device, out=buf[s])
                eng.ddpm_steps(mol, T - 1 - done, n, T, sched, noise=buf[:n])
                done += n
        flags = eng.read_flags()
        if flags & FLAG_CLAMPED:
            warnings.warn("Large molecule encountered in sampling")
        if flags & FLAG_NONFINITE:
            warnings.warn("Non-finite coordinates encountered in sampling")
        if flags & FLAG_CENTER:
            raise AssertionError("Center not at zero during sampling (assert_center_zero)")
        assert_center_zero(mol)
        return mol

    @torch.no_grad()
    def sample(self, batch_size):
        return self.p_sample_loop((batch_size, self.num_atoms, self.dims)) * self.norm_factor

    # ---- loss evaluation (the forward half of the training path: what trainer.eval_loss runs under no_grad, trainer.py:222-235)
    def q_mean_variance(self, x_start, t):
        mean = extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        return mean, extract(1.0 - self.alphas_cumprod, t, x_start.shape), extract(self.log_one_minus_alphas_cumprod, t, x_start.shape)

    @torch.no_grad()
    def assert_normal_kl(self, x_start, t, eps=1e-4):
        """KL(q(x_T | x_0) || N(0, I)) must vanish: enough diffusion steps (ddpm.py:173-193)."""
        assert_center_zero(x_start)
        mean1, _, logvar1 = self.q_mean_variance(x_start, t)
        logvar1 = logvar1.squeeze()
        kl = 0.5 * (-1.0 - logvar1 + torch.exp(logvar1) + (mean1 ** 2).sum(dim=(-2, -1)))
        assert kl.abs().max().item() <= eps, f"Normal KL check at T failed, max value: {kl.abs().max().item()}"

    def q_sample(self, x_start, t, noise=None):
        noise = center_zero(torch.randn_like(x_start) if noise is None else noise)
        return extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start \
            + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise

    @property
    def loss_fn(self):
        if self.loss_type == "l1":
            return F.l1_loss
        if self.loss_type == "l2":
            return F.mse_loss
        raise ValueError(f"invalid loss type {self.loss_type}")

    @torch.no_grad()
    def p_losses(self, x_start, t, noise=None):
        """The denoising loss at per-sample noise levels t [B] (ddpm.py:289-315).  The score network runs in the fused kernel with one t
        per sample (dff_score_dev_t); VALUE only: parameter gradients need a second-order reverse pass of the kernel (the training
        step, SURVEY.md 8f rank 4, is out of scope), so this is what `trainer.eval_loss` computes, not what `trainer.train` needs."""
        noise = center_zero(torch.randn_like(x_start) if noise is None else noise)
        x = center_zero(self.q_sample(x_start=x_start, t=t, noise=noise))
        model_out = center_zero(self.model(x, self.h, 1.0 * t / self.num_timesteps, alphas=None))
        loss = self.loss_fn(model_out, noise, reduction="none")
        return loss.reshape(loss.shape[0], -1).mean(dim=1).mean()

    def forward(self, mol, *args, t_diff_range=None, **kwargs):
        """Loss of a batch of structures in Angstrom at random noise levels (ddpm.py:317-337); evaluation only."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()) and self.training:
            raise DffError("GaussianDiffusion.forward in training mode needs parameter gradients (a second-order reverse pass of the "
                           "fused kernel): out of scope.  Call it in eval mode / under torch.no_grad() for the loss value (trainer.eval_loss)")
        mol = center_zero(mol) / self.norm_factor
        assert_center_zero(mol)
        b, n, d = mol.shape
        assert n == self.num_atoms and d == self.dims, f"Molecule shape must be {(self.num_atoms, self.dims)}"
        t = torch.multinomial(self.p2_loss_weight, b, replacement=True).long()
        self.assert_normal_kl(x_start=mol, t=torch.full((b,), self.num_timesteps - 1, device=mol.device, dtype=torch.long))
        return self.p_losses(mol, t, *args, **kwargs)
