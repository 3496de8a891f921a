"""cfg 3 end to end: N independent ALD reconstruction chains of one ACDC-shaped 4-coil R=40 acquisition, sharded over the
GPUs of one node (one process per GPU), posterior mean / std by ONE all-reduce of sufficient statistics, NRMSE / SSIM on the
device, result files in the reference's formats.  Random-init weights (no checkpoints offline): the images are meaningless,
the pipeline, its scaling and its timing are what this run shows.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_posterior.py --chains 105
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import parity_cases as C
from inverseproblemwithdiffusionmodel_b200 import chains as CH
from inverseproblemwithdiffusionmodel_b200.helpers import metrics as HM, results as HR

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=105)
ap.add_argument("--levels", type=int, default=2311, help="noise levels of the geometric 348 -> 0.01 schedule (acdc.yml: 2311)")
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--out", default="gpurun_out/posterior")
ap.add_argument("--seed", type=int, default=1234, help="one seed for every rank: the chains differ by their global ids")
args = ap.parse_args()
rank, local, world = CH.init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
n = args.size
cfg = C.make_config("ACDC", 128, n, args.levels, 348.0, device=str(dev))
torch.manual_seed(0)
net = C.NCSNv2Deepest(cfg).to(dev).eval()
A = C.SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 64, seed=0)
truth = C.phantom(1, 1, 1, n, n).to(dev)
y1 = A(truth)
sig = C.get_sigmas(cfg, mode="recons")
mine = CH.chain_partition(args.chains, world, rank)      # GLOBAL chain ids of this rank: chain i draws Philox(seed, chain i) on any rank
B = len(mine)
params = {"n_steps_each": 3, "step_lr": 9e-7, "denoise": True, "final_only": True}
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
stats = CH.PosteriorStats(n * n, dev)
if B > 0:                                                # a rank without chains (more ranks than chains) still joins the all-reduce
    sampler = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                              measurement=y1.repeat(1, B, 1, 1, 1), linear_tfm=A, seg=None, device=dev)
    recon = sampler(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", seed=args.seed, chain_ids=mine)[0]   # (B,1,H,W) on the host
    x = sampler.final_state                                                                                                          # same, on the device
    stats.add(x.reshape(B, n, n))
else:
    recon = torch.zeros(0, 1, n, n, dtype=torch.complex64)
    x = torch.zeros(0, 1, n, n, dtype=torch.complex64, device=dev)
post = stats.all_reduce().finalize((n, n))
torch.cuda.synchronize(); wall = time.perf_counter() - t0
t_all = torch.tensor([wall], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
if rank == 0:
    tm = truth.abs()
    rng = float(tm.max() - tm.min())
    m_mean = HM.compute_metrics(["NRMSE", "SSIM"], post["mag_mean"].reshape(1, 1, n, n), tm, data_range=rng)
    m_each = HM.compute_metrics(["NRMSE", "SSIM"], x.abs(), tm, reduce="mean", data_range=rng)
    HR.save_reconstruction(args.out, truth, y1, recon, linear_tfm=A, args_dict=vars(args))
    torch.save({k: v.cpu() for k, v in post.items() if k != "n"}, os.path.join(args.out, "posterior_stats.pt"))
    steps = args.levels * 3
    print(json.dumps({"config": f"cfg3: {args.chains} chains, ACDC-shaped {n}x{n}, 4 coils, R=40, {args.levels} levels x 3 steps + denoise",
                      "n_gpus": world, "chains_per_gpu_max": (args.chains + world - 1) // world, "posterior_chains": post["n"],
                      "wall_s_incl_graph_capture": round(float(t_all.item()), 2),
                      "chain_steps_per_s": round(args.chains * steps / float(t_all.item()), 1),
                      "nrmse_of_mean": float(m_mean["NRMSE"][0]), "ssim_of_mean": float(m_mean["SSIM"][0]),
                      "mean_nrmse_rank0_chains": float(m_each["NRMSE"]), "mean_ssim_rank0_chains": float(m_each["SSIM"]),
                      "mag_std_mean": float(post["mag_std"].mean()), "finite": bool(torch.isfinite(post["mag_mean"]).all()),
                      "files": sorted(os.listdir(args.out)), "weights": "random init (no checkpoints offline)"}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
