"""cfg 4 at the full schedule: ALD2DTime on a CINE127-shaped acquisition (1,24,1,128,128), 4 coils, live 24-frame mask,
cine127.yml's L = 1000 levels x 3 steps, once with the temporal TV step and once with the learned temporal prior
(NCSN3DShallow, cine127_1d-shaped schedule of 400 levels remapped onto the tail of the spatial one).  Random-init
weights: the run shows the whole pipeline end to end (graph switching per level, Philox noise, patch fold/unfold with
random rolls) staying finite over 3000 steps and what it costs, not image quality."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_cases as C
from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
dev = torch.device("cuda")
n, L = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cfg = C.make_config("CINE127", 128, n, L, 60.0, device="cuda")
torch.manual_seed(0)
net = C.NCSNv2Deepest(cfg).to(dev).eval()
sig = C.get_sigmas(cfg, mode="recons")
A = C.SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
truth = C.phantom(3, 24, 1, n, n).to(dev)
meas = A(truth).reshape(4, 1, 24, 1, n, n)
cfg_T = C.make_config("CINE127", 128, 24, 400, 40.0, device="cuda")
cfg_T.data.channels, cfg_T.data.channels_3d = 64, 1
torch.manual_seed(1)
net_T = NCSN3DShallow(cfg_T).to(dev).eval()
sig_T = C.get_sigmas(cfg_T)
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
for mode_T, lam in (("tv", 0.01), ("diffusion1d", 1.0)):
    if only and mode_T not in only:
        continue
    smp = C.ALD.ALD2DTime(C.L2Penalty(A), net_T, sig_T, (1, 24, 1, n, n), net, sig, {"n_steps_each": 3, "step_lr": 1e-4}, cfg,
                          measurement=meas, linear_tfm=A, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = smp(save_dir="/tmp", lr_scaled=1.0, mode_T=mode_T, lamda_T=lam, seed=2, if_random_shift=True)[0]
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    torch.set_grad_enabled(True)
    x = out.to(dev).reshape(24, 1, n, n)
    print(json.dumps({"config": f"cfg4 full schedule: (1,24,1,128,128), 4 coils, {L} levels x 3 steps, mode_T={mode_T}",
                      "wall_s_incl_graph_capture": round(wall, 2), "steps": 3 * L, "ms_per_step_avg": round(1e3 * wall / (3 * L), 2),
                      "finite": bool(torch.isfinite(x.abs()).all()), "max_abs": float(x.abs().max()),
                      "weights": "random init"}), flush=True)
