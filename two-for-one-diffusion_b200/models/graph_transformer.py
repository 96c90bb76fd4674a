"""`GraphTransformer` with the reference's constructor, state-dict layout and forward() contract
(models/graph_transformer.py:18-114), executed by the fused sm_100a kernel in libdff_b200.so.

Every constructor mode of the reference is supported: use_intrinsic_coords / use_distances (edge features x_j - x_i and
|x_j - x_i|^2) and use_abs_coords (x_i in the node input), conservative or not (graph_transformer.py:53-65, 99-100, 116-140).

The nn.Module tree below exists ONLY to own parameters under the reference's checkpoint key names
(`graphtransformer.layers.{l}.0.0.fn.to_q.weight`, ... SURVEY.md 3.4) so `load_state_dict` of a shipped
`model-best.pt["ema"]` works unchanged.  None of these sub-modules has a forward(): the arithmetic
(edge build, edge-conditioned attention, gated residual MLP, -dE/dx) lives in CUDA, and there is no CPU path.
"""
from typing import Optional

import torch
from torch import nn

from dff_b200 import DffError, ScoreEngine

HEADS, DIM_HEAD = 8, 64      # reference graph_transformer.py:213 (never overridden)


class _Params(nn.Module):
    """Parameter holder; calling it is an error by design."""

    def forward(self, *a, **k):
        raise DffError("this sub-module only stores parameters; call GraphTransformer.forward (CUDA)")


class Attention(_Params):
    def __init__(self, dim, edge_dim):
        super().__init__()
        inner = HEADS * DIM_HEAD
        self.to_q = nn.Linear(dim, inner)
        self.to_kv = nn.Linear(dim, 2 * inner)
        self.edges_to_kv = nn.Linear(edge_dim, inner)
        self.to_out = nn.Linear(inner, dim)


class PreNorm(_Params):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)


class GatedResidual(_Params):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Sequential(nn.Linear(3 * dim, 1, bias=False), nn.Sigmoid())


def _feed_forward(dim, mult=4):
    return nn.Sequential(nn.Linear(dim, dim * mult), nn.GELU(), nn.Linear(dim * mult, dim))


class GraphTransformerLucid(_Params):
    def __init__(self, dim, depth, edge_dim):
        super().__init__()
        self.layers = nn.ModuleList(
            nn.ModuleList([nn.ModuleList([PreNorm(dim, Attention(dim, edge_dim)), GatedResidual(dim)]),
                           nn.ModuleList([PreNorm(dim, _feed_forward(dim)), GatedResidual(dim)])])
            for _ in range(depth))


class GraphTransformer(nn.Module):
    def __init__(self, num_beads, hidden_nf, device="cpu", n_layers=4, use_intrinsic_coords: bool = False,
                 use_abs_coords: bool = True, use_distances: bool = True, conservative: bool = True):
        super().__init__()
        self.device = device
        self.num_beads, self.hidden_nf, self.n_layers = num_beads, hidden_nf, n_layers
        self.use_intrinsic_coords, self.use_distances = use_intrinsic_coords, use_distances
        self.use_abs_coords, self.conservative = use_abs_coords, conservative
        if conservative and not (use_intrinsic_coords or use_abs_coords or use_distances):
            raise DffError("a conservative network without intrinsic coordinates, distances or absolute coordinates does not depend "
                           "on x: the reference raises 'Gradient after computing forces is None' (graph_transformer.py:157-158)")
        in_node_nf = num_beads + 1 + use_abs_coords * 3                                                   # graph_transformer.py:53
        in_edge_nf = 3 * use_intrinsic_coords + use_distances + 1 * (not use_intrinsic_coords) * (not use_distances)     # :54-58
        self.node_embedding = nn.Linear(in_node_nf, hidden_nf)
        self.edge_embedding = nn.Linear(in_edge_nf, hidden_nf)
        self.node_decoder = nn.Linear(hidden_nf, 1 if conservative else 3)      # graph_transformer.py:62-65
        self.graphtransformer = GraphTransformerLucid(hidden_nf, n_layers, hidden_nf)
        self.max_batch = 4096
        self._eng, self._eng_key = None, None
        self.to(self.device)

    # ---- engine management: re-pack whenever parameters change (load_state_dict, .to(), optimizer step ...)
    def _param_key(self):
        ps = list(self.parameters())
        return (str(ps[0].device), tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps), self.max_batch)

    def engine(self, min_batch: int = 1) -> ScoreEngine:
        if min_batch > self.max_batch:
            self.max_batch = int(min_batch)
        key = self._param_key()
        if self._eng is None or key != self._eng_key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise DffError(f"GraphTransformer parameters are on {dev}; the score network only runs on a CUDA "
                               "device (B200, sm_100a) -- there is no CPU fallback")
            if self._eng is not None:
                self._eng.close()
            state = {k: v.detach() for k, v in self.state_dict().items()}
            self._eng = ScoreEngine(state, device=dev, max_batch=self.max_batch, use_intrinsic_coords=self.use_intrinsic_coords,
                                    use_distances=self.use_distances, use_abs_coords=self.use_abs_coords)
            self._eng_key = key
        return self._eng

    @staticmethod
    def _t_arg(t, batch, device):
        """One noise level for the batch (every caller on the sampling path: ddpm.py:240-247, langevin.py:77) -> python float;
        per-sample noise levels (graph_transformer.py:91 embeds t per sample; p_losses-style evaluation) -> [B] device tensor."""
        if not torch.is_tensor(t):
            return float(t)
        flat = t.reshape(-1)
        if flat.numel() not in (1, batch):
            raise DffError(f"t has {flat.numel()} entries for a batch of {batch}")
        if flat.numel() == 1 or bool((flat == flat[0]).all()):
            return float(flat[0])
        return flat.detach().to(device, torch.float32).contiguous()

    def _check_h(self, h):
        """`h` must be the bead one-hot matrix the node embedding was folded with (graph_transformer.py:99-103 with
        trainset.bead_onehot = eye(N)); any other feature matrix would silently be ignored."""
        if h is None:
            return
        if tuple(h.shape) != (self.num_beads, self.num_beads):
            raise DffError(f"h must be the [{self.num_beads},{self.num_beads}] bead one-hot matrix")
        key = (h.data_ptr(), h._version, str(h.device))
        if getattr(self, "_h_ok", None) != key:
            if not bool((h.detach().to("cpu", torch.float32) == torch.eye(self.num_beads)).all()):
                raise DffError("h must be the identity bead one-hot matrix (trainset.bead_onehot); other node features "
                               "are not supported by the fused kernel")
            self._h_ok = key

    def forward(self, x, h, t, return_energy=False, alphas: Optional[torch.Tensor] = None):
        """x [B,N,3]; h [N,N] bead one-hot (identity); t [B] / [B,1,1] = step/T (uniform or one value per sample).
        Returns forces (= -dE/dx, the epsilon prediction) [B,N,3], or energies [B,N,1] if return_energy.
        `alphas` is accepted and unused, exactly like the reference (graph_transformer.py:83)."""
        self._check_h(h)
        eng = self.engine(x.shape[0])
        t_shared = self._t_arg(t, x.shape[0], eng.device)
        xin = x.detach().to(eng.device, torch.float32).contiguous()
        if not self.conservative:       # the decoder output is the prediction; return_energy is ignored (:107-113)
            return eng.score(xin, t_shared, want_forces=True, want_energy=False)[0]
        eps, en = eng.score(xin, t_shared, want_forces=not return_energy, want_energy=return_energy)
        return en.unsqueeze(-1) if return_energy else eps
