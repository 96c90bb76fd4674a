"""`Langevin` integrator with the reference's constructor / simulate() contract (dynamics/langevin_cgnet.py:17-792),
driving the fused CUDA kernel: every save-interval chunk of BAOA(F)B or Brownian steps -- centre, force
evaluation (score forward + reverse mode), velocity/position/noise updates, kinetic energy -- is ONE launch.

Kept: BAOAB (friction given) and Brownian (friction None) schemes, save_interval / length checks, kinetic energies,
`simulated_coords` layout, progress logging (print or `{filename}_log.txt`), npy export every `export_interval` steps
(`{filename}_coords_NNN.npy`, `{filename}_kineticenergy_NNN.npy`, langevin_cgnet.py:568-603), resumable
`simulate(sub_interval)`; added: `state_dict()` / `load_state_dict()` for restarting an MD run in another process
(x, v, step counter, RNG state) -- SURVEY.md 8f rank 2.
Out of scope (unused by sample.py): save_forces/save_potential, `reference_beta` temperature ramps,
arbitrary (non-ForcesWrapper) models.

Noise: rng="torch" draws `torch.randn(size, generator=rng)` on the CPU once per step, in step order, exactly like the
reference (langevin_cgnet.py:469-472) and ships it to the GPU one chunk ahead of the kernel; rng="philox" draws
inside the kernel (no host traffic at all).
"""
import os
import time
import warnings

import numpy as np
import torch

from dff_b200 import DffError
from dff_b200 import _native as nat


class Langevin:
    def __init__(self, model, initial_coordinates, embeddings=None, dt=5e-4, beta=1.0, friction=None, masses=None,
                 diffusion=1.0, save_forces=False, save_potential=False, length=100, save_interval=10, random_seed=None,
                 device=torch.device("cpu"), export_interval=None, log_interval=None, log_type="write", filename=None,
                 rng="torch"):
        if save_forces or save_potential or embeddings is not None:
            raise DffError("save_forces / save_potential / embeddings are not part of the sampling path")
        if log_type not in ["print", "write"]:
            raise ValueError("log_type can be either 'print' or 'write'")
        if not hasattr(model, "model_gnn") or not hasattr(model, "force_scale"):
            raise DffError("the fused integrator needs a dynamics.langevin.ForcesWrapper force field")
        if getattr(model, "training", False):
            warnings.warn("model is in training mode; call model.eval() before simulating")
        self.model, self.initial_coordinates = model, initial_coordinates
        self.friction, self.masses = friction, masses
        if len(initial_coordinates.shape) != 3:
            raise ValueError("initial_coordinates shape must be [frames, beads, dimensions]")
        self.n_sims, self.n_beads, self.n_dims = initial_coordinates.shape
        self.length, self.save_interval = length, save_interval
        if length % save_interval != 0:
            raise ValueError("The save_interval must be a factor of the simulation length")
        self.dt, self.diffusion, self.beta = dt, diffusion, beta
        self.device, self.log_interval, self.log_type = torch.device(device), log_interval, log_type
        self.rng_mode = rng
        self.export_interval, self.filename = export_interval, filename
        # saving logs / numpys: same checks and file names as the reference (langevin_cgnet.py:352-398)
        if export_interval is not None and filename is None:
            raise RuntimeError("Must specify filename if export_interval isn't None")
        if log_interval is not None and log_type == "write" and filename is None:
            raise RuntimeError("Must specify filename if log_interval isn't None and log_type=='write'")
        if export_interval is not None:
            if length // export_interval >= 1000:
                raise ValueError("Simulation saving is not implemented if more than 1000 files will be generated")
            if os.path.isfile("{}_coords_000.npy".format(filename)):
                raise ValueError("{} already exists; choose a different filename.".format("{}_coords_000.npy".format(filename)))
            if export_interval % save_interval != 0:
                raise ValueError("Numpy saving must occur at a multiple of save_interval")
            self._npy_file_index = 0
        if log_interval is not None:
            if log_interval % save_interval != 0:
                raise ValueError("Logging must occur at a multiple of save_interval")
            if log_type == "write":
                self._log_file = filename + "_log.txt"
                if os.path.isfile(self._log_file):
                    raise ValueError("{} already exists; choose a different filename.".format(self._log_file))
        if friction is not None:
            if masses is None:
                raise RuntimeError("if friction is not None, masses must be given")
            if len(masses) != self.n_beads:
                raise ValueError("mass list length must be number of CG beads")
            self.vscale = np.exp(-dt * friction)
            self.noisescale = np.sqrt(1 - self.vscale * self.vscale)
            if diffusion != 1:
                warnings.warn("Diffusion other than 1. was provided, but Langevin dynamics do not use it")
        else:
            self._dtau = diffusion * dt
            if masses is not None:
                warnings.warn("Masses were provided, but will not be used since friction is None (i.e., infinte).")
        self.kinetic_energies = [] if friction is not None else None
        self.rng = torch.default_generator if random_seed is None else torch.Generator().manual_seed(random_seed)
        self.random_seed = random_seed
        self._simulated = False
        self._chunk_noise = None

    # ------------------------------------------------------------------
    def _params(self) -> "nat.MdParams":
        p = nat.MdParams()
        p.integrator = nat.DFF_MD_BROWNIAN if self.friction is None else nat.DFF_MD_BAOAB
        p.t_norm = float(self.model.t_int) / float(self.model.diffusion_steps)
        p.force_scale = self.model.force_scale()
        p.dt, p.beta = float(self.dt), float(self.beta)
        if self.friction is None:
            p.dtau = float(self._dtau)
        else:
            p.vscale, p.noisescale = float(self.vscale), float(self.noisescale)
        return p

    def _log(self, msg):
        if self.log_interval is None:
            return
        if self.log_type == "print":
            print(msg)
        else:                                   # langevin_cgnet.py:544-557
            with open(self._log_file, "a") as f:
                f.write(msg + "\n")

    def _save_numpy(self, frames_d, ke_d, first, last):
        """frames [first, last) of this simulate() call -> {filename}_coords_NNN.npy ([n_sims, frames, beads, 3]) and, with
        friction, {filename}_kineticenergy_NNN.npy ([n_sims, frames]); numbering continues across simulate() calls."""
        key = "{:03d}".format(self._npy_file_index)
        np.save("{}_coords_{}.npy".format(self.filename, key), frames_d[first:last].permute(1, 0, 2, 3).contiguous().cpu().numpy())
        if ke_d is not None:
            np.save("{}_kineticenergy_{}.npy".format(self.filename, key), ke_d[first:last].t().contiguous().cpu().numpy())
        self._npy_file_index += 1

    # ---- restart (not in the reference, which can only resume inside one process: langevin_cgnet.py:719-722)
    def state_dict(self):
        """Everything needed to continue this run elsewhere: positions, velocities, step counter and the RNG stream."""
        if not hasattr(self, "x_old"):
            raise DffError("nothing to save yet: call simulate() first")
        return {"x": self.x_old.detach().cpu().clone(), "v": None if self.v_old is None else self.v_old.detach().cpu().clone(),
                "t": int(self.t), "seed": int(self._seed), "rng_mode": self.rng_mode, "rng_state": self.rng.get_state().clone(),
                "npy_file_index": getattr(self, "_npy_file_index", 0), "length": int(self.length),
                "save_interval": int(self.save_interval)}

    def load_state_dict(self, state):
        if state["rng_mode"] != self.rng_mode:
            raise DffError("restart state was written with rng=%r, this integrator uses rng=%r" % (state["rng_mode"], self.rng_mode))
        if tuple(state["x"].shape) != (self.n_sims, self.n_beads, self.n_dims):
            raise ValueError("restart coordinates have shape %s, expected %s" % (tuple(state["x"].shape), (self.n_sims, self.n_beads, self.n_dims)))
        dev = self.model.model_gnn.engine(self.n_sims).device
        self.x_old = state["x"].to(dev, torch.float32).contiguous().clone()
        self.v_old = None if state["v"] is None else state["v"].to(dev, torch.float32).contiguous().clone()
        if (self.v_old is None) != (self.friction is None):
            raise DffError("restart state and integrator disagree on the scheme (BAOAB needs velocities, Brownian has none)")
        self.t, self._seed = int(state["t"]), int(state["seed"])
        if self.rng is torch.default_generator:
            torch.set_rng_state(state["rng_state"])
        else:
            self.rng.set_state(state["rng_state"])
        if self.export_interval is not None:
            self._npy_file_index = int(state.get("npy_file_index", 0))

    def _draw_chunk(self, n_steps, pinned):
        """n_steps CPU draws, one torch.randn per step like the reference, into a pinned staging buffer."""
        for s in range(n_steps):
            torch.randn(size=(self.n_sims, self.n_beads, self.n_dims), generator=self.rng, out=pinned[s])
        return pinned

    def simulate(self, sub_interval=None, reference_beta=None):
        """Returns simulated_coords as numpy [n_sims, n_frames, n_beads, 3] (frames are x_new, not re-centred)."""
        if reference_beta is not None:
            raise DffError("temperature ramps (reference_beta) are not part of the sampling path")
        if self.device.type != "cuda":
            raise DffError("Langevin.simulate needs a CUDA device; there is no CPU fallback")
        sub_interval = self.length if sub_interval is None else sub_interval
        if sub_interval % self.save_interval != 0:
            raise ValueError("The save_interval must be a factor of the simulated interval")
        B, N, si = self.n_sims, self.n_beads, self.save_interval
        done_before = getattr(self, "t", 0)
        # frames this call can still produce: the loop stops at self.length, so never allocate (and return) more than that
        n_save = min(sub_interval, max(self.length - done_before, 0)) // si
        if n_save == 0 and hasattr(self, "x_old"):          # the run is complete: nothing to integrate, no RNG draws
            self.simulated_coords = np.zeros((B, 0, N, 3), dtype=np.float32)
            if self.friction is not None:
                self.kinetic_energies = np.zeros((B, 0), dtype=np.float32)
            return self.simulated_coords
        eng = self.model.model_gnn.engine(B)
        dev = eng.device
        if not hasattr(self, "x_old"):
            self._log("Generating {} simulations of length {} saved at {}-step intervals ({})".format(
                B, self.length, si, time.asctime()))
            self.x_old = self.initial_coordinates.detach().to(dev, torch.float32).contiguous().clone()
            self.v_old = torch.zeros_like(self.x_old) if self.friction is not None else None
            self.t = 0
            self._seed = int(torch.randint(0, 2 ** 62, (1,), generator=self.rng).item()) if self.rng_mode == "philox" else 0
        prm = self._params()
        mass = torch.tensor(self.masses if self.masses is not None else [1.0] * N, dtype=torch.float32, device=dev)
        frames_d = torch.empty(n_save, B, N, 3, device=dev)
        ke_d = torch.empty(n_save, B, device=dev) if self.friction is not None else None
        coords_h = torch.empty(n_save, B, N, 3).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        use_host_rng = self.rng_mode != "philox"
        if use_host_rng and n_save > 0:
            stage = [torch.empty(si, B, N, 3).pin_memory() for _ in range(2)]
            noise_d = [torch.empty(si, B, N, 3, device=dev) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            self._draw_chunk(si, stage[0])
            with torch.cuda.stream(copy_stream):
                noise_d[0].copy_(stage[0], non_blocking=True)
                ready[0].record()
        if self.t == 0 and self.model.norm is None and self.log_interval is not None:
            self.model(self.x_old - self.x_old.mean(dim=1, keepdim=True))     # prints "Forces (norm)" like langevin.py:89-91
        done_chunks, npy_start = 0, 0
        main = torch.cuda.current_stream(dev)
        while self.t < self.length and done_chunks < n_save:
            k = done_chunks
            cur = k & 1
            if use_host_rng:
                main.wait_event(ready[cur])
            eng.langevin_steps(self.x_old, self.v_old, si, prm, mass, noise=noise_d[cur] if use_host_rng else None,
                               seed=self._seed, offset=self.t, save_interval=si, frames=frames_d[k:k + 1],
                               ke=None if ke_d is None else ke_d[k:k + 1])
            if use_host_rng:
                consumed[cur].record(main)
                if k + 1 < n_save:          # draw + upload the next chunk's noise while the GPU integrates this one
                    nxt = cur ^ 1
                    if k >= 1:
                        consumed[nxt].synchronize()
                    self._draw_chunk(si, stage[nxt])
                    with torch.cuda.stream(copy_stream):
                        noise_d[nxt].copy_(stage[nxt], non_blocking=True)
                        ready[nxt].record()
            self.t += si
            done_chunks += 1
            if self.export_interval is not None and (done_chunks * si) % self.export_interval == 0:
                self._save_numpy(frames_d, ke_d, npy_start, done_chunks)
                npy_start = done_chunks
            if self.log_interval is not None and (self.t % self.log_interval) == 0:
                self._log("{}/{} time points saved ({})".format(self.t // si, self.length // si, time.asctime()))
        if self.export_interval is not None and npy_start < done_chunks:          # the remainder (langevin_cgnet.py:775-778)
            self._save_numpy(frames_d, ke_d, npy_start, done_chunks)
        coords_h.copy_(frames_d, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        flags = eng.read_flags()
        if flags & nat.FLAG_NONFINITE:
            warnings.warn("Non-finite coordinates encountered in the simulation")
        self.simulated_coords = coords_h.permute(1, 0, 2, 3).contiguous().numpy()
        if ke_d is not None:
            self.kinetic_energies = ke_d.t().contiguous().cpu().numpy()
        self.simulated_forces, self.simulated_potential = None, None
        self._simulated = True
        return self.simulated_coords
