#!/bin/bash
# GPU box: compute-sanitizer passes over one small force evaluation + 2 fused BAOAB steps of the tcgen05 kernel, for both attention
# flavours (DFF_ATTN) and a distance + absolute-coordinate network (HMMA path with the distance channel)
cat > /tmp/san_case.py <<'PY'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, os.path.join(ROOT, "two-for-one-diffusion_b200"), os.path.join(ROOT, "tests")]
import torch
from helpers import net_params
from dff_b200 import ScoreEngine, _native as nat
from oracle.weights import synthetic_net_params
mol, N = sys.argv[1], int(sys.argv[2])
if mol == "modes":
    eng = ScoreEngine(synthetic_net_params(N, 64, 2, 5, in_edge=1, in_node_extra=3), max_batch=8, use_intrinsic_coords=False, use_distances=True, use_abs_coords=True)
else:
    eng = ScoreEngine(net_params(mol), max_batch=8)
x = torch.randn(3, N, 3, generator=torch.Generator().manual_seed(1)).cuda()
x = (x - x.mean(1, keepdim=True)).contiguous()
eps, en = eng.score(x, 0.02, want_energy=True)
prm = nat.MdParams(nat.DFF_MD_BAOAB, 0.02, -0.01, 7e-4, 0.9992, 0.0394, 0.0343, 0.0)
v = torch.zeros_like(x)
eng.langevin_steps(x, v, 2, prm, torch.full((N,), 12.0, device="cuda"), noise=None, seed=3, offset=0)
torch.cuda.synchronize()
print(mol, eng.last_config, os.environ.get("DFF_ATTN", "default"), float(eps.abs().max()), float(x.abs().max()))
PY
for tool in memcheck racecheck synccheck; do
  for case in "chignolin 10 simt" "chignolin 10 mma" "trp_cage 20 mma" "modes 12 mma"; do
    set -- $case
    echo "=== compute-sanitizer --tool $tool : $1 N=$2 attention=$3"
    DFF_ATTN=$3 timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=dff_fused_tc python /tmp/san_case.py $1 $2 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error|tc |chignolin|trp_cage|modes" | head -8
  done
done
