"""ORACLE (test infrastructure only) -- literal CPU restatement of the score network.

Follows /root/reference/models/graph_transformer.py operation by operation, as
plain functions over a parameter dict (the checkpoint's `ema_model.model.*`
tensors with that prefix stripped), so that the numbers it produces are the
reference's numbers and its cost on a CPU is the reference's cost:

  score_forward            GraphTransformer.forward            graph_transformer.py:77-114
  _edge_features           GraphTransformer.get_edge_attr      graph_transformer.py:116-140
  _attention               Attention.forward                   graph_transformer.py:229-258
  _gated_residual          GatedResidual.forward               graph_transformer.py:202-205
  _transformer_stack       GraphTransformerLucid.forward       graph_transformer.py:318-329
  forces by autograd       compute_forces                      graph_transformer.py:143-159

It deliberately materialises the [B,N,N,H] edge embedding and the three
[8B,N,N,64] tensors per layer exactly like the reference does (that is where
the reference spends its time).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

HEADS = 8          # graph_transformer.py:213 (never overridden at :67-73)
DIM_HEAD = 64      # graph_transformer.py:213
LN_EPS = 1e-5      # nn.LayerNorm default, graph_transformer.py:182


def strip_prefix(state: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}


def net_params_from_ema(ema_state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`model-best.pt["ema"]` -> score-network tensors (sample.py:154-167, :179 uses ema_model only)."""
    return strip_prefix(ema_state, "ema_model.model.")


def num_layers(p: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"graphtransformer.layers.{n}.0.0.norm.weight" in p:
        n += 1
    return n


def infer_config(p: Dict[str, torch.Tensor]) -> dict:
    """Recover (N, H, L, edge/node mode) from tensor shapes (graph_transformer.py:53-65)."""
    H, in_node = p["node_embedding.weight"].shape
    in_edge = p["edge_embedding.weight"].shape[1]
    n_out = p["node_decoder.weight"].shape[0]
    return dict(hidden=H, in_node=in_node, in_edge=in_edge, layers=num_layers(p),
                conservative=(n_out == 1))


def _center(x: torch.Tensor) -> torch.Tensor:
    # utils.py:65-70
    return x - x.mean(dim=1, keepdim=True)


def _edge_features(x: torch.Tensor, use_intrinsic: bool, use_dist: bool) -> torch.Tensor:
    # graph_transformer.py:116-140; diff[b,i,j] = x[b,j] - x[b,i]
    B, N, _ = x.shape
    if not use_intrinsic and not use_dist:
        return torch.zeros(B, N, N, 1, dtype=x.dtype, device=x.device)
    diff = x[:, None, :, :] - x[:, :, None, :]
    if use_intrinsic and not use_dist:
        return diff
    sq = (diff * diff).sum(dim=3, keepdim=True)
    if use_dist and not use_intrinsic:
        return sq
    return torch.cat([diff, sq], dim=3)


def _split_heads(t: torch.Tensor) -> torch.Tensor:
    # einops 'b ... (h d) -> (b h) ... d', h=8  (graph_transformer.py:237-239)
    B = t.shape[0]
    mid = t.shape[1:-1]
    t = t.reshape(B, *mid, HEADS, DIM_HEAD)
    t = torch.movedim(t, -2, 1)                       # b h ... d
    return t.reshape(B * HEADS, *mid, DIM_HEAD)


def _attention(p, pre: str, nodes_n: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    q = F.linear(nodes_n, p[pre + "to_q.weight"], p[pre + "to_q.bias"])
    kv = F.linear(nodes_n, p[pre + "to_kv.weight"], p[pre + "to_kv.bias"])
    k, v = kv.chunk(2, dim=-1)
    e_kv = F.linear(edges, p[pre + "edges_to_kv.weight"], p[pre + "edges_to_kv.bias"])
    q, k, v, e_kv = map(_split_heads, (q, k, v, e_kv))
    k = k[:, None, :, :] + e_kv                       # [8B,N,N,64]   :241-245
    v = v[:, None, :, :] + e_kv
    sim = torch.einsum("bid,bijd->bij", q, k) * (DIM_HEAD ** -0.5)
    # mask is all-True on this path (:104, :249-253) -> masked_fill is a no-op
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bijd->bid", attn, v)
    BH, N, D = out.shape
    out = out.reshape(BH // HEADS, HEADS, N, D).permute(0, 2, 1, 3).reshape(BH // HEADS, N, HEADS * D)
    return F.linear(out, p[pre + "to_out.weight"], p[pre + "to_out.bias"])


def _gated_residual(w: torch.Tensor, x: torch.Tensor, res: torch.Tensor) -> torch.Tensor:
    gate = torch.sigmoid(F.linear(torch.cat((x, res, x - res), dim=-1), w))
    return x * gate + res * (1 - gate)


def _transformer_stack(p, nodes: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    H = nodes.shape[-1]
    for l in range(num_layers(p)):
        a = f"graphtransformer.layers.{l}.0."
        f = f"graphtransformer.layers.{l}.1."
        n_hat = F.layer_norm(nodes, (H,), p[a + "0.norm.weight"], p[a + "0.norm.bias"], LN_EPS)
        att = _attention(p, a + "0.fn.", n_hat, edges)
        nodes = _gated_residual(p[a + "1.proj.0.weight"], att, nodes)
        m_hat = F.layer_norm(nodes, (H,), p[f + "0.norm.weight"], p[f + "0.norm.bias"], LN_EPS)
        hid = F.gelu(F.linear(m_hat, p[f + "0.fn.0.weight"], p[f + "0.fn.0.bias"]))
        ff = F.linear(hid, p[f + "0.fn.2.weight"], p[f + "0.fn.2.bias"])
        nodes = _gated_residual(p[f + "1.proj.0.weight"], ff, nodes)
    return nodes


def score_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, t_norm, *,
                  use_intrinsic_coords: bool = True, use_abs_coords: bool = False,
                  use_distances: bool = False, return_energy: bool = False,
                  h: Optional[torch.Tensor] = None) -> torch.Tensor:
    """== GraphTransformer.forward(x, h, t, return_energy) in eval mode.

    x [B,N,3]; t_norm scalar / [B] / [B,1,1] (diffusion index / T); h defaults to eye(N)
    (ddpm.py:42 + dataset_utils_empty.py:218).  Returns forces [B,N,3]
    (= -d sum(E) / dx for conservative nets) or per-bead energies [B,N,1].
    """
    conservative = p["node_decoder.weight"].shape[0] == 1
    x = _center(x.detach())
    B, N, _ = x.shape
    dtype = x.dtype
    if h is None:
        h = torch.eye(N, dtype=dtype, device=x.device)
    t = torch.as_tensor(t_norm, dtype=dtype).to(x.device).reshape(-1, 1, 1)
    if t.shape[0] == 1:
        t = t.expand(B, 1, 1)
    t = t.repeat(1, N, 1)
    hb = h.to(x.device, dtype).unsqueeze(0).repeat(B, 1, 1)
    with torch.enable_grad() if conservative else torch.no_grad():
        if conservative:
            x = x.requires_grad_(True)
        edges = _edge_features(x, use_intrinsic_coords, use_distances)
        edges = F.linear(edges, p["edge_embedding.weight"], p["edge_embedding.bias"])
        node_in = torch.cat((hb, x, t), dim=2) if use_abs_coords else torch.cat((hb, t), dim=2)
        nodes = F.linear(node_in, p["node_embedding.weight"], p["node_embedding.bias"])
        nodes = _transformer_stack(p, nodes, edges)
        out = F.linear(nodes, p["node_decoder.weight"], p["node_decoder.bias"])
        if not conservative:
            return out.detach()
        if return_energy:
            return out.detach()
        (grad,) = torch.autograd.grad(out, x, grad_outputs=torch.ones_like(out))
    return -grad


def to_dtype(p: Dict[str, torch.Tensor], dtype) -> Dict[str, torch.Tensor]:
    return {k: v.to(dtype) for k, v in p.items()}


def literal_flops_per_sample(N: int, H: int, L: int) -> float:
    """SURVEY.md 8d 'reference-literal' FLOPs (fwd + x-backward, GEMM terms, 2 FLOP/MAC)."""
    I = HEADS * DIM_HEAD
    fwd = 2 * N * N * 3 * H + 2 * N * (N + 1) * H + L * (
        2 * N * H * I + 2 * N * H * 2 * I + 2 * N * N * H * I + 4 * N * N * I
        + 2 * N * I * H + 16 * N * H * H + 12 * N * H) + 2 * N * H
    bwd = fwd - 2 * N * H * 3 * I - 2 * N * (N + 1) * H
    return float(fwd + bwd)


def collapsed_flops_per_sample(N: int, H: int, L: int) -> float:
    """SURVEY.md 8d 'collapsed' FLOPs -- what the CUDA kernel executes."""
    I = HEADS * DIM_HEAD
    per_layer = (2 * N * H * 3 * I + 2 * (2 * N * I * 3) + 2 * 2 * N * N * (I + 24)
                 + 2 * N * I * H + 16 * N * H * H + 12 * N * H)
    fwd = L * per_layer + 2 * N * H
    bwd = fwd - 2 * N * H * 3 * I
    return float(fwd + bwd)
