"""Sampling CLI -- same flags, checkpoint format and outputs as the reference's sample.py (:18-98, :101-249):

    python sample.py --model_path saved_models/chignolin --gen_mode iid      --num_samples_eval 1000
    python sample.py --model_path saved_models/chignolin --gen_mode langevin --parallel_sim 256 --n_timesteps 10000

reads {model_path}/args.pickle + model-{ckpt}.pt["ema"], writes
{model_path}/main_eval_output_{gen_mode}[_{append}]/sample-{gen_mode}.pt (CPU float32 [n, N, 3], Angstrom) and a
.pdb of the first 1000 frames.  Differences: the score network and both samplers run in the fused sm_100a kernel;
multi-GPU is one process per GPU (`torchrun --nproc-per-node 8 sample.py ...`) with the batch sharded across ranks
and ONE all-gather of the sampled coordinates at the end, instead of nn.DataParallel; mdtraj / ema_pytorch /
tensorboard are not needed.  Extra flags: --rng {torch,philox}, --seed, --out_dir.
"""
import argparse
import os
import pickle
import sys
import time
from os.path import join
from pathlib import Path

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

from datasets.dataset_utils_empty import get_dataset          # noqa: E402
from dff_b200.ema import EMA                                   # noqa: E402
from dff_b200.pdb import save_pdb                              # noqa: E402
from dynamics.langevin import LangevinDiffusion, temp_dict    # noqa: E402
from evaluate.evaluators import sample_from_model             # noqa: E402
from models import get_model                                   # noqa: E402
from models.ddpm import GaussianDiffusion                      # noqa: E402
from utils import SamplerWrapper                               # noqa: E402


def build_parser():
    p = argparse.ArgumentParser(description="coarse-graining-evaluator")
    p.add_argument("--model_path", type=str, required=True, help="root directory where models and args are stored")
    p.add_argument("--model_checkpoint", type=str, default="best", help="best, last, 1, 2, 3, ...")
    p.add_argument("--gen_mode", type=str, default="iid", help="generative mode, either iid or langevin")
    p.add_argument("--append_exp_name", type=str, default=None,
                   help="append this text to the results/main_eval_output folder name, append only gen_mode if None (default)")
    p.add_argument("--data_folder", type=str, default=None, help="directory root where data is stored (must be None here)")
    # i.i.d. generation
    p.add_argument("--num_samples_eval", type=int, default=1000, help="number of samples for i.i.d. generation")
    p.add_argument("--batch_size_gen", type=int, default=256, help="batch size for evaluation")
    # Langevin simulation
    p.add_argument("--masses", type=eval, default=None, help="Units in g/mol")
    p.add_argument("--friction", type=float, default=1, help="friction, usually 1")
    p.add_argument("--parallel_sim", type=int, default=100, help="Number of parallel simulations")
    p.add_argument("--n_timesteps", type=int, default=10000, help="number of timesteps")
    p.add_argument("--save_interval", type=int, default=250, help="save interval (in timesteps)")
    p.add_argument("--noise_level", type=int, default=20, help="diffusion model noise level for extracting force fields")
    p.add_argument("--dt", type=float, default=None, help="time step in ps; None = derived from the diffusion model")
    p.add_argument("--temp_data", type=float, default=None, help="temperature in Kelvin.")
    p.add_argument("--temp_sim", type=float, default=None, help="temperature in Kelvin")
    p.add_argument("--kb", type=str, default="consistent", help="consistent, kcal")
    # additions
    p.add_argument("--rng", type=str, default="torch", choices=["torch", "philox"],
                   help="torch: the reference's torch.randn stream (default); philox: in-kernel RNG, no per-step host traffic")
    p.add_argument("--seed", type=int, default=None, help="torch.manual_seed(seed + rank) before sampling")
    p.add_argument("--out_dir", type=str, default=None, help="write outputs here instead of inside model_path")
    return p


def _dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend)
    return world, rank, local


def gather_samples(local: torch.Tensor, world: int, rank: int):
    """One all-gather of the sampled coordinates (rank-major = simulation-major, matching langevin.py:209-211)."""
    if world == 1:
        return local
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = local.to(dev).contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat(parts, dim=0).cpu()


def main(samp_args):
    world, rank, local = _dist_setup()
    if not torch.cuda.is_available():
        raise SystemExit("sample.py needs a CUDA device (B200, sm_100a): the score network has no CPU fallback")
    torch.cuda.set_device(local)
    device = "cuda"
    if samp_args.seed is not None:
        torch.manual_seed(samp_args.seed + rank)
    with open(join(samp_args.model_path, "args.pickle"), "rb") as f:
        args = pickle.load(f)
    if samp_args.temp_data is None:
        samp_args.temp_data = temp_dict[args.mol.upper()]
    if samp_args.temp_sim is None:
        samp_args.temp_sim = temp_dict[args.mol.upper()]
    suffix = f"_{samp_args.gen_mode}" + ("" if samp_args.append_exp_name is None else f"_{samp_args.append_exp_name}")
    eval_folder = Path(join(samp_args.out_dir or samp_args.model_path, "main_eval_output" + suffix))
    if rank == 0:
        eval_folder.mkdir(exist_ok=True, parents=samp_args.out_dir is not None)
    args.data_folder = samp_args.data_folder
    trainset, _, _ = get_dataset(args.mol, args.mean0, args.data_folder, args.fold,
                                 shuffle_before_splitting=args.shuffle_data_before_splitting)
    norm_factor = trainset.std if args.scale_data else 1.0
    model_nn = get_model(args, trainset, device)
    if rank == 0:
        print(model_nn)
    ddpm = GaussianDiffusion(model=model_nn, features=trainset.bead_onehot, num_atoms=trainset.num_beads,
                             timesteps=args.diffusion_steps, norm_factor=norm_factor, loss_weights=args.loss_weights,
                             rng=samp_args.rng).to(device)
    model = EMA(ddpm)
    data_dict = torch.load(samp_args.model_path + f"/model-{samp_args.model_checkpoint}.pt", map_location="cpu")
    model.load_state_dict(data_dict["ema"])
    out = generate_samples(model, trainset, samp_args.noise_level, args, device, eval_folder, samp_args, world, rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return out


def generate_samples(model, trainset, noise_level, args, device, eval_folder, samp_args, world=1, rank=0):
    sampler = SamplerWrapper(model.ema_model).to(device).eval()
    t0 = time.time()
    if samp_args.gen_mode == "iid":
        local_mol = sample_from_model(sampler, samp_args.num_samples_eval // world, max(1, samp_args.batch_size_gen // world),
                                      verbose=rank == 0)
    elif samp_args.gen_mode == "langevin":
        if rank == 0:
            print("Total number of samples to save using Langevin Dynamics: "
                  f"{int(samp_args.parallel_sim * samp_args.n_timesteps / samp_args.save_interval)}")
        # initial states are drawn from the model itself (reference sample.py:197-214)
        init_mol = sample_from_model(sampler, samp_args.parallel_sim // world, max(1, samp_args.batch_size_gen // world),
                                     verbose=rank == 0)
        masses = samp_args.masses
        if masses is None:
            masses = [12.8 if "alanine" in args.mol else 12.0] * trainset.num_beads
        sim = LangevinDiffusion(model.ema_model, init_mol, samp_args.n_timesteps, save_interval=samp_args.save_interval,
                                t=noise_level, diffusion_steps=args.diffusion_steps, temp_data=samp_args.temp_data,
                                temp_sim=samp_args.temp_sim, dt=samp_args.dt, masses=masses, friction=samp_args.friction,
                                kb=samp_args.kb, rng=samp_args.rng)
        local_mol = sim.sample()
    else:
        raise Exception("Wrong argument 'gen_mode'")
    sampled_mol = gather_samples(local_mol, world, rank)
    if rank == 0:
        print(f"{len(sampled_mol)} structures in {time.time() - t0:.1f} s")
        torch.save(sampled_mol, str(eval_folder) + f"/sample-{samp_args.gen_mode}.pt")
        save_pdb(str(eval_folder) + f"/sample-{samp_args.gen_mode}.pdb", sampled_mol[0:1000].numpy(), trainset.topology)
    return sampled_mol


if __name__ == "__main__":
    main(build_parser().parse_args())
