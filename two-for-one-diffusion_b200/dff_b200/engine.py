"""ScoreEngine -- owns one `dff_model_t` handle and exposes the three device entry points on torch
tensors (device pointers in, device pointers out).  This is plumbing: all arithmetic happens in
libdff_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _native as nat

_GLOBAL_KEYS = ("node_embedding.weight", "node_embedding.bias", "edge_embedding.weight", "edge_embedding.bias",
                "node_decoder.weight", "node_decoder.bias")
_LAYER_KEYS = ("0.0.norm.weight", "0.0.norm.bias", "0.0.fn.to_q.weight", "0.0.fn.to_q.bias", "0.0.fn.to_kv.weight",
               "0.0.fn.to_kv.bias", "0.0.fn.edges_to_kv.weight", "0.0.fn.edges_to_kv.bias", "0.0.fn.to_out.weight",
               "0.0.fn.to_out.bias", "0.1.proj.0.weight", "1.0.norm.weight", "1.0.norm.bias", "1.0.fn.0.weight",
               "1.0.fn.0.bias", "1.0.fn.2.weight", "1.0.fn.2.bias", "1.1.proj.0.weight")
SCHED_KEYS = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
              "posterior_mean_coef2", "posterior_log_variance_clipped")


def ordered_weight_names(n_layers: int):
    names = list(_GLOBAL_KEYS)
    for l in range(n_layers):
        names += [f"graphtransformer.layers.{l}.{k}" for k in _LAYER_KEYS]
    return names


def count_layers(state: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"graphtransformer.layers.{n}.0.0.norm.weight" in state:
        n += 1
    return n


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
        raise nat.DffError(f"expected a contiguous float32 tensor on {device}, got {t.dtype} on {t.device}")
    return t


class ScoreEngine:
    """Device-resident packed model.  `state` holds the score-network tensors under the reference's
    state-dict names (`GraphTransformer.state_dict()`, i.e. checkpoint keys minus `ema_model.model.`)."""

    def __init__(self, state: Dict[str, torch.Tensor], device="cuda:0", max_batch: int = 4096, use_intrinsic_coords: bool = True,
                 use_distances: bool = False, use_abs_coords: bool = False):
        L = count_layers(state)
        H, in_node = state["node_embedding.weight"].shape
        n_out = state["node_decoder.weight"].shape[0]
        in_edge = state["edge_embedding.weight"].shape[1]
        want_edge = 3 * bool(use_intrinsic_coords) + bool(use_distances) + (not use_intrinsic_coords and not use_distances)
        if in_edge != want_edge or n_out not in (1, 3):
            raise nat.DffError(f"edge_embedding has {in_edge} input features, but use_intrinsic_coords={use_intrinsic_coords}, "
                               f"use_distances={use_distances} needs {want_edge} (graph_transformer.py:54-58); "
                               "pass the flags the network was trained with (args.pickle)")
        self.conservative = n_out == 1
        self.use_intrinsic_coords, self.use_distances, self.use_abs_coords = bool(use_intrinsic_coords), bool(use_distances), bool(use_abs_coords)
        self.num_beads, self.hidden, self.n_layers = in_node - 1 - (3 if use_abs_coords else 0), H, L
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise nat.DffError("ScoreEngine needs a CUDA device; there is no CPU path")
        self.max_batch = int(max_batch)
        host = [state[k].detach().to("cpu", torch.float32).contiguous() for k in ordered_weight_names(L)]
        arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        opts = nat.ModelOpts(int(self.conservative), int(self.use_intrinsic_coords), int(self.use_distances), int(self.use_abs_coords))
        nat.check(nat.lib().dff_model_create_v2(C.byref(h), idx, self.num_beads, H, L, arr, len(host), self.max_batch, C.byref(opts)))
        self._h = h
        self._flags = torch.zeros(1, dtype=torch.int32, device=self.device)

    def close(self):
        if getattr(self, "_h", None):
            nat.lib().dff_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- queries
    @property
    def launches(self) -> int:
        return int(nat.lib().dff_model_launch_count(self._h))

    @property
    def last_config(self) -> str:
        """Launch configuration of the last call: "tc" (tcgen05/TMEM kernel) or "wide"/"tall"/"duo" (mma.sync kernel)."""
        return nat.lib().dff_model_last_config(self._h).decode()

    @property
    def flops_per_sample(self) -> float:
        return float(nat.lib().dff_model_flops_per_sample(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- == GraphTransformer.forward
    def score(self, x: torch.Tensor, t_norm, want_forces=True, want_energy=False):
        """t_norm: a python float (one noise level for the batch) or a [B] float32 device tensor (one per sample)."""
        x = _dev_f32(x, self.device)
        B = x.shape[0]
        assert x.shape[1:] == (self.num_beads, 3), x.shape
        eps = torch.empty_like(x) if want_forces else None
        en = torch.empty(B, self.num_beads, device=self.device, dtype=torch.float32) if want_energy else None
        with torch.cuda.device(self.device):
            if torch.is_tensor(t_norm):
                t = _dev_f32(t_norm.reshape(-1), self.device)
                if t.numel() != B:
                    raise nat.DffError(f"t has {t.numel()} entries for a batch of {B}")
                nat.check(nat.lib().dff_score_dev_t(self._h, _ptr(x), _ptr(t), B, _ptr(eps), _ptr(en), self._stream()))
            else:
                nat.check(nat.lib().dff_score_dev(self._h, _ptr(x), float(t_norm), B, _ptr(eps), _ptr(en), self._stream()))
        return eps, en

    # ---- == GaussianDiffusion.p_sample_loop slice (in place on x)
    def ddpm_steps(self, x: torch.Tensor, t_start: int, n_steps: int, T: int, sched: Sequence[torch.Tensor],
                   noise: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0) -> int:
        x = _dev_f32(x, self.device)
        sp = (C.c_void_p * 5)(*[_dev_f32(s, self.device).data_ptr() for s in sched])
        if noise is not None:
            noise = _dev_f32(noise, self.device)
            assert noise.shape == (n_steps,) + tuple(x.shape), noise.shape
        with torch.cuda.device(self.device):
            nat.check(nat.lib().dff_ddpm_steps_dev(self._h, _ptr(x), x.shape[0], int(t_start), int(n_steps), int(T), sp,
                                                   _ptr(noise), seed, offset, _ptr(self._flags), self._stream()))
        return 0

    # ---- == Langevin.simulate slice (in place on x, v)
    def langevin_steps(self, x, v, n_steps: int, prm: "nat.MdParams", mass: torch.Tensor, noise=None, seed=0,
                       offset=0, save_interval=0, frames=None, ke=None):
        x = _dev_f32(x, self.device)
        if v is not None:
            v = _dev_f32(v, self.device)
        if noise is not None:
            noise = _dev_f32(noise, self.device)
            assert noise.shape == (n_steps,) + tuple(x.shape), noise.shape
        with torch.cuda.device(self.device):
            nat.check(nat.lib().dff_langevin_steps_dev(self._h, _ptr(x), _ptr(v), x.shape[0], int(n_steps), C.byref(prm),
                                                       _ptr(mass), _ptr(noise), seed, offset, int(save_interval),
                                                       _ptr(frames), _ptr(ke), _ptr(self._flags), self._stream()))

    def read_flags(self, reset=True) -> int:
        f = int(self._flags.item())
        if reset:
            self._flags.zero_()
        return f

    # ---- test hook
    def debug_stash(self):
        rows, samples, npad = C.c_int(), C.c_int(), C.c_int()
        lf = C.c_int64()
        offs = (C.c_int64 * 11)()
        nat.check(nat.lib().dff_debug_stash_layout(self._h, C.byref(rows), C.byref(samples), C.byref(npad), C.byref(lf), offs))
        n = lf.value * self.n_layers + rows.value * self.hidden
        buf = torch.empty(n, dtype=torch.float32)
        got = nat.lib().dff_debug_read_stash(self._h, C.c_void_p(buf.data_ptr()), n)
        if got < 0:
            nat.check(int(got))
        return dict(rows=rows.value, samples=samples.value, npad=npad.value, layer_floats=lf.value,
                    offsets=list(offs), data=buf[:got])
