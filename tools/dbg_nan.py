import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_cases as C
dev = torch.device("cuda", 0)
n, B, levels = 256, 14, int(sys.argv[1]) if len(sys.argv) > 1 else 24
cfg = C.make_config("ACDC", 128, n, levels, 348.0, device=str(dev))
torch.manual_seed(0)
net = C.NCSNv2Deepest(cfg).to(dev).eval()
A = C.SENSE("exp", 4, 40, 1 / 64, (1, n, n), 0)
A.random_under_fourier.mask = C.keep_center_mask(n, 40, 1 / 64, seed=0)
y1 = A(C.phantom(1, 1, 1, n, n).to(dev))
sig = C.get_sigmas(cfg, mode="recons")
params = {"n_steps_each": 3, "step_lr": 9e-7, "denoise": True, "final_only": True}
s = C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sig, params, cfg,
                                    measurement=y1.repeat(1, B, 1, 1, 1), linear_tfm=A, seg=None, device=dev)
fc = s(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", seed=1234, return_chain=True)
for k in range(levels * 3):
    fc["step"]()
    torch.cuda.synchronize()
    st, g = fc["state"], fc["grad"]
    print(k, "sigma %.4g" % float(sig[k // 3]), "state max %.4g" % float(st.abs().max()), "grad max %.4g" % float(g.abs().max()),
          "finite", bool(torch.isfinite(st).all()), bool(torch.isfinite(g).all()), flush=True)
    if not torch.isfinite(st).all():
        break
