"""Developer aid (GPU box): compare every stashed intermediate of the CUDA forward pass with
oracle/collapsed_ref.py, region by region, to localise a numerical bug."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "two-for-one-diffusion_b200"), os.path.join(ROOT, "tests")]
import torch
from helpers import net_params, load, rel_err
from oracle import collapsed_ref, score_ref
from dff_b200 import ScoreEngine

mol = sys.argv[1] if len(sys.argv) > 1 else "chignolin"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = net_params(mol)
c = load(f"score_{mol}.pt")["cases"][0]
x = c["x"][:B].contiguous()
eng = ScoreEngine(p, max_batch=64)
eps, en = eng.score(x.cuda(), c["t_norm"], want_energy=True)
torch.cuda.synchronize()
f64, e64, saved = collapsed_ref.forward_backward(score_ref.to_dtype(p, torch.float64), x.double(), c["t_norm"], want_stash=True)
print("forces rel err", rel_err(eps, f64), " energy rel err", rel_err(en, e64), "flags")
st = eng.debug_stash()
R, S, NP, LF, off, data = st["rows"], st["samples"], st["npad"], st["layer_floats"], st["offsets"], st["data"]
N, H = eng.num_beads, eng.hidden
rows = min(B, S) * N      # CTA 0 holds the first S samples
print("R", R, "S", S, "NP", NP)
nb = rows // N
xc = (x.double() - x.double().mean(1, keepdim=True))[:nb]
saved = [{k: (v[:nb] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B and k != "Ah" else v) for k, v in s_.items()} for s_ in saved]
for l, s in enumerate(saved):
    base = l * LF
    def reg(i, n): return data[base + off[i]: base + off[i] + n]
    nin = reg(0, R * H).view(R, H)[:rows]
    print(f"L{l} n_in ", rel_err(nin, s["n_in"].reshape(rows, H)))
    st1 = reg(1, 2 * R).view(R, 2)[:rows]
    print(f"L{l} rstd1", rel_err(st1[:, 1], s["r1"].reshape(rows)))
    qkv = reg(2, 8 * R * 192).view(8, R, 192)[:, :rows]
    q = s["q"].reshape(rows, 8, 64).permute(1, 0, 2)
    Ah = s["Ah"]
    e = torch.einsum("hdc,rc->hrd", Ah, xc.reshape(rows, 3))
    k = s["k"].reshape(rows, 8, 64).permute(1, 0, 2) + e
    v = s["v"].reshape(rows, 8, 64).permute(1, 0, 2) + e
    print(f"L{l} q    ", rel_err(qkv[:, :, :64], q), " k'", rel_err(qkv[:, :, 64:128], k), " v'", rel_err(qkv[:, :, 128:], v))
    pp = reg(3, 8 * R * NP).view(8, R, NP)[:, :rows, :N]
    pr = s["p"].permute(1, 0, 2, 3).reshape(8, rows, N)
    print(f"L{l} p    ", rel_err(pp, pr))
    print(f"L{l} att  ", rel_err(reg(4, R * H).view(R, H)[:rows], s["att"].reshape(rows, H)))
    print(f"L{l} g1   ", rel_err(reg(5, R)[:rows], s["g1"].reshape(rows)))
    print(f"L{l} m    ", rel_err(reg(6, R * H).view(R, H)[:rows], s["m"].reshape(rows, H)))
    print(f"L{l} h1   ", rel_err(reg(8, 4 * R * H).view(R, 4 * H)[:rows], s["h1"].reshape(rows, 4 * H)))
    print(f"L{l} ff   ", rel_err(reg(9, R * H).view(R, H)[:rows], s["ff"].reshape(rows, H)))
    print(f"L{l} g2   ", rel_err(reg(10, R)[:rows], s["g2"].reshape(rows)))
