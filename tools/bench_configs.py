"""Throughput of the other BASELINE.json configurations (bench.py's line is cfg 2):
  cfg 1: NCSNv2 unconditional ALD, (16,1,28,28), mnist.yml schedule (L=232, 5 steps each, step_lr 6.2e-6)
  cfg 4: ALD2DTime on a CINE127-shaped volume (1,24,1,128,128), 4 coils, live 24-frame mask, mode_T none / tv /
         diffusion1d (NCSN3DShallow temporal prior on 2*256 patches of 8x8x24, every level running the temporal step)
Each reports steps/s of the captured step graph (device-resident state), CUDA-event timed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_cases as C
dev = torch.device("cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def time_steps(step, warm=3, n=20):
    for _ in range(warm):
        step()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- cfg 1
cfg = C.make_config("MNIST", 128, 28, 232, 50.0, device="cuda")
torch.manual_seed(0)
net = C.NCSNv2(cfg).to(dev).eval()
sig = C.get_sigmas(cfg)
params = {"n_steps_each": 5, "step_lr": 6.2e-6, "denoise": True, "final_only": True}
s = C.ALD.ALDUnconditionalSampler((16, 1, 28, 28), net, sig, params, cfg, device=dev)
import time
t0 = time.perf_counter(); out = s(seed=1)[0]; torch.cuda.synchronize(); wall = time.perf_counter() - t0
fc = list(s._fast_cache.values())[0]
ms = time_steps(fc["step"])
print(json.dumps({"config": "cfg1 MNIST 28x28 NCSNv2 ngf128 batch 16", "ms_per_step": round(ms, 4), "steps_per_s": round(1e3 / ms, 1),
                  "images_per_s": round(16e3 / ms, 1), "full_chain_1160_steps_wall_s_incl_capture": round(wall, 2),
                  "finite": bool(torch.isfinite(out).all()), "launches_per_step": s.launches_per_step}), flush=True)
del s, net, fc
torch.set_grad_enabled(True)
# ---- cfg 4
n = 128
cfg = C.make_config("CINE127", 128, n, 1000, 60.0, device="cuda")
torch.manual_seed(0)
net = C.NCSNv2Deepest(cfg).to(dev).eval()
sig = C.get_sigmas(cfg, mode="recons")
A = C.SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)
meas = A(C.phantom(3, 24, 1, n, n).to(dev)).reshape(4, 1, 24, 1, n, n)
for mode_T in ("none", "tv"):
    smp = C.ALD.ALD2DTime(C.L2Penalty(A), None, sig[-10:], (1, 24, 1, n, n), net, sig[:12], {"n_steps_each": 3, "step_lr": 1e-4}, cfg,
                          measurement=meas, linear_tfm=A, device=dev)
    out = smp(save_dir="/tmp", lr_scaled=1.0, mode_T=mode_T, lamda_T=0.01, seed=2)[0]
    fc = list(smp._fast_cache.values())[0]
    ms = time_steps(fc["step"])
    print(json.dumps({"config": f"cfg4 CINE127-shaped (1,24,1,128,128) 4 coils, mode_T={mode_T}", "ms_per_step": round(ms, 4),
                      "steps_per_s": round(1e3 / ms, 1), "frame_forwards_per_step": 48, "conv_tflops": round(48 * 209.59e9 / ms / 1e9, 1),
                      "finite": bool(torch.isfinite(out.abs()).all()), "launches_per_step": smp.launches_per_step}), flush=True)
    torch.set_grad_enabled(True)

# ---- cfg 4 with the learned temporal prior
from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
cfg_T = C.make_config("CINE127", 128, 24, 400, 40.0, device="cuda")
cfg_T.data.channels, cfg_T.data.channels_3d = 64, 1
torch.manual_seed(1)
net_T = NCSN3DShallow(cfg_T).to(dev).eval()
sig_T = C.get_sigmas(cfg_T)
smp = C.ALD.ALD2DTime(C.L2Penalty(A), net_T, sig_T, (1, 24, 1, n, n), net, sig[-12:], {"n_steps_each": 3, "step_lr": 1e-4}, cfg,
                      measurement=meas, linear_tfm=A, device=dev)
out = smp(save_dir="/tmp", lr_scaled=1.0, mode_T="diffusion1d", lamda_T=1.0, seed=2)[0]
fc = list(smp._fast_cache.values())[0]
ms_T = time_steps(fc["step_T"])
ms_0 = time_steps(fc["step"])
# temporal net alone: 512 patches, one forward
vol = torch.rand(512, 8, 24, 8, device=dev); gv = torch.empty_like(vol); lab = torch.zeros(512, dtype=torch.long, device=dev)
ms_net = time_steps(lambda: net_T.forward_into(vol, lab, gv))
# 3x3x3 convolutions of NCSN3DShallow: 19 x 128->128, 2 x 128->256, 2 x 256->256 at T; 27 x 256->256 at T/2
flop_T = 512 * 8 * 8 * 2 * 27 * (24 * (19 * 128 * 128 + 2 * 128 * 256 + 2 * 256 * 256) + 12 * 27 * 256 * 256)
print(json.dumps({"config": "cfg4 CINE127-shaped (1,24,1,128,128) 4 coils, mode_T=diffusion1d (NCSN3DShallow ngf128, 512 patches 8x8x24)",
                  "ms_per_step_with_temporal": round(ms_T, 4), "ms_per_step_without": round(ms_0, 4), "steps_per_s": round(1e3 / ms_T, 1),
                  "temporal_net_forward_ms": round(ms_net, 4), "temporal_net_conv_tflops": round(flop_T / ms_net / 1e9, 1),
                  "finite": bool(torch.isfinite(out.abs()).all()), "launches_per_step": smp.launches_per_step}), flush=True)
