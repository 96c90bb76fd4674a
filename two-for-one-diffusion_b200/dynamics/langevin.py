"""Diffusion model -> force field -> Langevin dynamics, with the reference's API (dynamics/langevin.py:46-212).

`ForcesWrapper` is the force field  F = -eps_theta(x, t*) / (kbt_inv * sqrt(1 - abar_t*));  `LangevinDiffusion`
does the unit bookkeeping and owns a `Langevin` integrator, whose loop runs inside the fused CUDA kernel.
"""
import torch
from torch import nn

from dynamics.langevin_cgnet import Langevin

KBOLTZMANN = 1.38064852e-23
AVOGADRO = 6.022140857e23
JPERKCAL = 4184
KB = 0.83144626181      # Boltzmann constant in g/mol * A^2 / ps^2 / K  (reference langevin.py:9)

_T300 = ("alanine_dipeptide_fuberlin".upper(), "alanine_dipeptide_mdshare".upper())
temp_dict = {**{k: 300 for k in _T300}, "CHIGNOLIN": 340, "TRP_CAGE": 290, "BBA": 325, "VILLIN": 360, "WW_DOMAIN": 360,
             "NTL9": 355, "BBL": 298, "PROTEIN_B": 340, "HOMEODOMAIN": 360, "PROTEIN_G": 350, "ALPHA3D": 370,
             "LAMBDA_REPRESSOR": 350}
temp_dict_pt = {**{k: 450 for k in _T300}, **{k: 500 for k in temp_dict if k not in _T300}}


class ForcesWrapper(nn.Module):
    """model(x, embeddings) -> (potential placeholder [B], forces [B,N,3])   (reference langevin.py:46-92)."""

    def __init__(self, model_diff, t=10, diffusion_steps=1000, kbt_inv=1.0):
        super().__init__()
        self.model_gnn = model_diff.model.eval()
        self.t = torch.Tensor([t]).to(model_diff.device)
        self.t_int, self.diffusion_steps = int(t), int(diffusion_steps)
        self.sqrt_one_minus_alphas_cumprod = model_diff.sqrt_one_minus_alphas_cumprod[t]
        self.sqrt_alphas_cumprod = model_diff.sqrt_alphas_cumprod
        self.t_norm = self.t / float(diffusion_steps)
        self.kbt_inv = kbt_inv
        self.one_hot = model_diff.h
        self.norm = None

    def force_scale(self) -> float:
        """forces = force_scale * eps_theta; the scalar the fused integrator applies in-kernel."""
        return -1.0 / (float(self.kbt_inv) * float(self.sqrt_one_minus_alphas_cumprod))

    def forward(self, x_old, embeddings=None):
        eps = self.model_gnn(x_old, self.one_hot, self.t_norm)
        forces = -eps / self.kbt_inv / self.sqrt_one_minus_alphas_cumprod
        if self.norm is None:
            self.norm = torch.mean(torch.norm(forces.cpu(), dim=2))
            print(f"Forces (norm) {self.norm}")
        return torch.zeros(x_old.shape[0]), forces


class LangevinDiffusion:
    """Reference langevin.py:95-212: same arguments, same printed diagnostics, same output layout
    ([n_sims * n_frames, N, 3] in Angstrom, simulation-major)."""

    def __init__(self, model_diff, init_mol, n_timesteps=1000000, save_interval=250, t=15, diffusion_steps=1000,
                 temp_data=300, temp_sim=300, dt=2e-3, masses=[12.8] * 5, friction=1, kb="consistent",
                 exchange_interval=5000, rng="torch", random_seed=None):
        print(f"norm factor:{model_diff.norm_factor}")
        self.norm_factor = model_diff.norm_factor
        self.device = model_diff.device
        self.one_minus_alphas_cumprod = 1 - model_diff.alphas_cumprod[t].item()
        if kb == "consistent":
            self.kb_inv = 1 / KB * self.norm_factor ** 2
        elif kb == "kcal":
            self.kb_inv = JPERKCAL / KBOLTZMANN / AVOGADRO * (self.norm_factor ** 2) / 100
        else:
            raise Exception("Wrong kb value")
        self.model_forces = ForcesWrapper(model_diff, t, diffusion_steps, kbt_inv=self.kb_inv / temp_data)
        gamma = 1 if friction is None else friction
        diffusion_constant = 1 / masses[0] if friction is None else 1
        if dt is None:       # step size matched to the diffusion noise level (langevin.py:160-168)
            dt = self.one_minus_alphas_cumprod * gamma * masses[0] * self.kb_inv / temp_data
        self.sim = Langevin(self.model_forces, init_mol / self.norm_factor, length=n_timesteps, save_interval=save_interval,
                            beta=self.kb_inv / temp_sim, save_potential=False, device=self.device, log_interval=save_interval,
                            log_type="print", diffusion=diffusion_constant, masses=masses, friction=friction, dt=dt,
                            rng=rng, random_seed=random_seed)
        print(f"Diffusion model Beta : {model_diff.betas[t]}")
        print(f"Diffusion model sqrt_alphas_cumprod {model_diff.sqrt_alphas_cumprod[t]}")
        print(f"Diffusion model sqrt_one_minus_alphas_cumprod {model_diff.sqrt_one_minus_alphas_cumprod[t]}")
        print(f"Diffusion model one_minus_alphas_cumprod {self.one_minus_alphas_cumprod}")
        print(f"dt*kb*T/M/gamma: {dt * temp_data / self.kb_inv / masses[0] / gamma} (should be on a similar scale as one_minus_alphas_cumprod)")
        print(f"dt: {dt: .8f} (ps)")
        print(f"KbT: {temp_data/self.kb_inv: .4f}")

    def sample(self):
        traj = torch.Tensor(self.sim.simulate())                       # [n_sims, n_frames, N, 3]
        return traj.reshape(-1, traj.size(2), traj.size(3)) * self.norm_factor
