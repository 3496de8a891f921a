#!/usr/bin/env python
"""bench.py -- ALD chain-steps/s on the cfg-2 workload of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (configs[1]): ACDC-shaped 2-D complex MRI, 256x256, 4-coil SENSE, keep-centre column mask
with Bernoulli rate 1/R (R = 40, center_lines_frac = 1/64), NCSNv2Deepest (acdc.yml, ngf 128) with
default random init, step_lr 9e-7, L2Penalty data consistency with lr_scaled 1e6, seg guidance off.
One "step" = the body of ALDInvSegProximalRealImag's inner loop for every chain of the rank: one
batched score forward over the real and imaginary planes of all chains (2 * chains images), the
fused Langevin update + SENSE proximal kernel, the schedule advance -- replayed as one CUDA graph.
`value` = chains * steps / time with the chain state resident in HBM; `e2e` = the same metric
through the public sampler call with the measurement in pinned host memory and the result read
back to the host, on a short schedule (config.e2e_call_steps ALD steps per call).
Chains shard over ranks with no data-path collective (weak scaling); the only collective is the
posterior mean/std all-reduce after the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "ALD chain-steps/sec"
UNIT = "chain-steps/s"
CONV_FLOP_PER_FORWARD_256 = 838.36e9      # NCSNv2Deepest ngf 128, one 256x256 image


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chains", type=int, default=14, help="chains per GPU (105 chains over 8 GPUs -> 14)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--coils", type=int, default=4)
    ap.add_argument("--R", type=float, default=40.0)
    ap.add_argument("--center-frac", type=float, default=1 / 64)
    ap.add_argument("--e2e-levels", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, chains, line_count):
    return {"workload": f"cfg2: ACDC-shaped {args.size}x{args.size} complex, {args.coils}-coil SENSE, R={args.R:g} "
                        f"(center_lines_frac={args.center_frac:.5f}: {line_count} of {args.size} lines sampled), "
                        f"NCSNv2Deepest ngf128, ALD + L2Penalty prox",
            "chains_per_gpu": chains, "images_per_forward": 2 * chains, "n_steps_each": 3, "step_lr": 9e-7, "lr_scaled": 1e6,
            "schedule_levels": 2311, "operand_dtype": "f16 (fp32 accumulate, fp32 residual streams)",
            "l2": "activations (>= 0.9 GB per tensor) exceed the 126 MB L2; no flush needed",
            "e2e_call_steps": args.e2e_levels * 3}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ oracle legs
def oracle_step_time(args, n_warm, n_timed):
    """Seconds per cfg-2 ALD chain-step (one chain) of the oracle port on the host cores."""
    import torch
    from oracle import mri_ops as M, scorenet as SN, ald as OALD
    import parity_cases as C
    torch.set_num_threads(os.cpu_count())
    n = args.size
    cfg = C.make_config("ACDC", 128, n, 2311, 348.0, device="cpu")
    torch.manual_seed(0)
    net = C.NCSNv2Deepest(cfg)          # parameter container only (default init); the oracle does the arithmetic
    Pd = {k: v.detach() for k, v in net.state_dict().items()}
    maps = M.exp_coil_maps(args.coils, n, n, 0)
    mask = M.keep_center_mask(n, args.R, args.center_frac, seed=0)
    y = M.sense_forward(C.phantom(1, 1, 1, n, n), maps, mask)
    sig = Pd["sigmas"]
    x = M.sense_adjoint(y, maps)
    xr, xi = x.real, x.imag
    labels = torch.zeros(1, dtype=torch.long)
    step = 9e-7 * (sig[0] / sig[-1]) ** 2
    times = []
    with torch.no_grad():
        for it in range(n_warm + n_timed):
            t0 = time.perf_counter()
            gr = SN.score_forward("NCSNv2Deepest", Pd, xr, labels)
            gi = SN.score_forward("NCSNv2Deepest", Pd, xi, labels)
            xr = OALD.langevin_update(xr, gr, torch.randn_like(xr), step)
            xi = OALD.langevin_update(xi, gi, torch.randn_like(xi), step)
            z = M.l2_prox_sense_closed_form(xr + 1j * xi, y, maps, mask, 9e-7 * 1e6, 1.0)
            xr, xi = z.real, z.imag
            dt = time.perf_counter() - t0
            if it >= n_warm:
                times.append(dt)
    return sum(times) / len(times), int(mask.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_timed = max(1, min(args.steps, 6))       # ~8 s of CPU per step: keep the arm within a few minutes
    n_warm = 1 if args.warmup > 0 else 0
    sec, lines = oracle_step_time(args, n_warm, n_timed)
    val = 1.0 / sec
    cfg = workload_config(args, 1, lines)
    cfg["chains_per_gpu"] = 1
    cfg["images_per_forward"] = 2
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"{n_timed} timed cfg-2 ALD steps of one chain (2 score forwards at {args.size}^2 + update + prox), "
                                      f"torch CPU with {os.cpu_count()} threads, after {n_warm} warm-up"},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ native arm
def sense_point(torch, C, _lib, L, dev, hbm, coils, size, batch, R, frac):
    """SENSE forward / adjoint / fused ALD step at one cfg-5 sweep point: algorithmic GB/s (fwd/adj
    8N(1+Nc) + 4*Nc*H*W, fused step 32N + 4*Nc*H*W bytes, N = B*H*W) over CUDA-event time, L2 flushed between
    iterations, best of 5.  `adjoint` is A^H on masked data (mask applied); `conj_op_unmasked` is the reference's
    SENSE.conj_op signature, which transforms every column (quirk Q3)."""
    A = C.SENSE("exp", coils, R, frac, (1, size, size), 0)
    A.random_under_fourier.mask = C.keep_center_mask(size, R, frac, seed=0)
    x = torch.randn(batch, 1, size, size, dtype=torch.complex64, device=dev)
    S = A(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def best(fn):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts)

    N = batch * size * size
    b_fa = 8 * N * (1 + coils) + 4 * coils * size * size
    b_st = 32 * N + 4 * coils * size * size
    state = torch.randn(2, batch, size, size, device=dev); grad = torch.randn_like(state); bvec = torch.randn_like(state)
    mre, _ = A.device_maps(dev); m, frames = A.device_mask(dev)
    sc = _lib.AldScalars(0.1, 0.4, 0.01, 1.0)
    plan = A.device_plan(dev, size)
    step = lambda: _lib.check(L.ipdm_ald_sense_step_plan(plan.handle, state.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), mre.data_ptr(), None,
                                                         coils, batch, size, sc, None, None, _lib.rng(1, 0), _lib.stream()))
    t = {"forward": best(lambda: A(x)), "adjoint": best(lambda: A.conj_op_masked(S)), "conj_op_unmasked": best(lambda: A.conj_op(S)),
         "fused_ald_step": best(step)}
    out = {"point": f"{coils} coils, {size}x{size}, batch {batch}, R={R} ({int(A.random_under_fourier.mask.sum())} lines), k-space {8 * coils * N / 1e6:.0f} MB",
           "hbm_peak_gbs": hbm, "l2": "256 MB flush between iterations"}
    for k, ms in t.items():
        byt = b_st if k == "fused_ald_step" else b_fa
        out[k] = {"ms": ms, "gbs": byt / ms / 1e6, "frac": byt / ms / 1e6 / hbm}
    return out


def run_native(args):
    import torch
    import torch.distributed as dist
    import parity_cases as C
    from inverseproblemwithdiffusionmodel_b200 import _lib, chains as CH
    rank, local, world = CH.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    n, B = args.size, args.chains
    cfg = C.make_config("ACDC", 128, n, 2311, 348.0, device=str(dev))
    torch.manual_seed(0)
    net = C.NCSNv2Deepest(cfg).to(dev).eval()
    A = C.SENSE("exp", args.coils, args.R, args.center_frac, (1, n, n), 0)
    A.random_under_fourier.mask = C.keep_center_mask(n, args.R, args.center_frac, seed=0)
    lines = int(A.random_under_fourier.mask.sum())
    y1 = A(C.phantom(1, 1, 1, n, n).to(dev))                       # (Nc,1,1,H,W)
    meas_host = y1.repeat(1, B, 1, 1, 1).cpu().pin_memory()
    sig = C.get_sigmas(cfg, mode="recons")
    params = {"n_steps_each": 3, "step_lr": 9e-7, "denoise": True, "final_only": True}
    my_chains = CH.chain_partition(B * world, world, rank)           # global chain ids of this rank (weak scaling)
    seed = 1234

    def make_sampler(sigmas, measurement):
        return C.ALD.ALDInvSegProximalRealImag(C.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sigmas, params, cfg,
                                               measurement=measurement, linear_tfm=A, seg=None, device=dev)
    kw = dict(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", seed=seed + rank)

    # ---- device-resident throughput: K replays of the captured step -------------------------------
    sampler = make_sampler(sig, meas_host.to(dev))
    chain = sampler(return_chain=True, **kw)
    step = chain["step"]
    launches_per_step = step.launches
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = B * world * args.steps / (ms_total / 1e3)
    finite = bool(torch.isfinite(chain["state"]).all())

    # ---- the path's one collective: posterior statistics over all chains ---------------------------
    stats = CH.PosteriorStats(n * n, dev)
    xc = torch.complex(chain["state"][0], chain["state"][1]).reshape(B, 1, n, n)
    stats.add(xc)
    stats.all_reduce()
    post = stats.finalize((n, n))

    # ---- end to end through the public call: host measurement in, host result out ------------------
    import numpy as np
    e2e_sig = torch.tensor(np.exp(np.linspace(np.log(348.0), np.log(0.01), args.e2e_levels))).float().to(dev)
    s2 = make_sampler(e2e_sig, meas_host)
    s2.params = dict(params, denoise=False)
    s2(**kw)                                                     # first call captures the graph
    torch.cuda.synchronize()
    n_calls = max(1, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n_calls):
        res = s2(**kw)
    e1.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    call_steps = args.e2e_levels * 3
    e2e_value = B * world * call_steps * n_calls / (float(t2.item()) / 1e3)
    h2d = meas_host.numel() * 8
    d2h = res[0].numel() * 8
    torch.set_grad_enabled(True)

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    # By launch-list share (profiles/r01_launch_breakdown.txt) the dominant kernel is k_conv_halo<7,1>: the
    # 128->128 3x3 convolution at 256^2 with residual + f32 + f16 stores (47 of 165 launches, 36 % of the step);
    # the f16-store-only variant of the same shape (30 launches) is timed beside it.  Same tensors as in the
    # step: 2*chains images, far larger than L2.
    import ctypes
    N = 2 * B
    x16 = torch.randn(N, n, n, 128, device=dev).half()
    w16 = (torch.randn(128, 9, 128, device=dev) / 34).half()
    o16 = torch.empty_like(x16)
    o32 = torch.empty(N, n, n, 128, device=dev)
    res = torch.randn(N, n, n, 128, device=dev)
    conv_flop = 2.0 * N * n * n * 128 * 128 * 9

    def time_conv(desc, reps=10):
        for _ in range(3):
            _lib.check(L.ipdm_conv_igemm(ctypes.byref(desc), _lib.stream()), "igemm")
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            _lib.check(L.ipdm_conv_igemm(ctypes.byref(desc), _lib.stream()), "igemm")
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    d_res = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, res.data_ptr(), o32.data_ptr(), o16.data_ptr(), None,
                          N, n, n, 128, 128, 9, 1, _lib.CONV_F16_ELU)
    d_f16 = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, None, None, o16.data_ptr(), None, N, n, n, 128, 128, 9, 1, _lib.CONV_F16_ELU)
    ms_res, ms_f16 = time_conv(d_res), time_conv(d_f16)
    burst, sustained, hbm, peak_kind = measured_peaks()
    achieved = conv_flop / (ms_res / 1e3) / 1e12
    step_conv_tflops = 2 * B * CONV_FLOP_PER_FORWARD_256 * (n / 256) ** 2 / (ms_total / args.steps / 1e3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this very shape
    # (profiles/r01_ncu_conv_halo_res_mode.txt: 1.412 GB read + 1.359 GB written at 28 images of 256^2; algorithmic 2.82 GB:
    # a little of the f16 input is still L2-resident from the previous launch)
    traffic = 2.770720e9 if (N == 28 and n == 256) else None
    alg_bytes = N * n * n * 128 * (2 + 4 + 4 + 2)
    # Which roofline bounds this variant?  Arithmetic intensity = 2*128*1152 FLOP / 1536 B per pixel = 192 FLOP/B, below
    # the ridge of the measured peaks (burst tensor / HBM copy ~ 258 FLOP/B): by the roofline model it is HBM-bound,
    # so `achieved`/`peak` are GB/s; the tensor-pipe view of the same launch and of the store-only variant
    # (AI 576 FLOP/B, tensor-bound) are reported beside it.
    ai = conv_flop / alg_bytes
    ridge = burst * 1e12 / (hbm * 1e9)
    tensor_view = {"achieved_tflops": achieved, "peak_tflops": burst, "frac": achieved / burst}
    gbs = alg_bytes / ms_res / 1e6
    if ai < ridge:
        head = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}
    else:
        head = {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst}
    roofline = dict(head, **{
                "kernel": "k_conv_halo<res+f32+f16> (128->128 3x3 @%dx%d, %d images)" % (n, n, N), "traffic": traffic,
                "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge, "tensor_view": tensor_view,
                "peak_source": f"MEASURED_PEAKS.json HBM copy GB/s and bf16 burst TFLOP/s ({peak_kind}); f16 kind::f16 has the same nominal rate as bf16",
                "ms_per_launch": ms_res, "flop_per_launch": conv_flop, "algorithmic_bytes_per_launch": alg_bytes,
                "same_shape_f16_store_only": {"bound": "tensor", "ms_per_launch": ms_f16, "achieved": conv_flop / (ms_f16 / 1e3) / 1e12,
                                              "peak": burst, "unit": "TFLOP/s", "frac": conv_flop / (ms_f16 / 1e3) / 1e12 / burst},
                "whole_step_conv_tflops": step_conv_tflops, "whole_step_frac_of_sustained": step_conv_tflops / sustained})

    # ---- SENSE operator GB/s (second headline metric) at a cfg-5 point whose k-space exceeds L2 ------------
    sense = None
    if rank == 0:
        del x16, o16, o32, res
        sense = sense_point(torch, C, _lib, L, dev, hbm, coils=args.coils, size=n, batch=64, R=args.R, frac=args.center_frac)
        if world == 1:   # the largest cfg-5 sweep point (k-space 4.3 GB >> L2): where SURVEY 8(d) evaluates the HBM fraction
            torch.cuda.empty_cache()
            sense["largest_sweep_point"] = sense_point(torch, C, _lib, L, dev, hbm, coils=32, size=512, batch=64, R=40.0, frac=1 / 64)
            torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, _ = oracle_step_time(args, 1, 2)
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"2 timed cfg-2 ALD steps of ONE chain (2 score forwards at {n}^2 + update + prox) after 1 warm-up, "
                         f"oracle port on torch CPU with {os.cpu_count()} threads"}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16", "data": "synthetic", "config": workload_config(args, B, lines),
           "clocks": clock_info,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "note": f"one bench step = one public sampler call of {call_steps} ALD steps; {n_calls} calls timed"},
           "gpu_launches": launches_per_step * args.steps,
           "launches_per_step": launches_per_step,
           "roofline": roofline, "sense": sense, "cpu_baseline": cpu,
           "state_finite": finite, "posterior_chains": post["n"], "my_chains": len(my_chains)}
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
