#!/usr/bin/env python
"""bench.py -- ALD chain-steps/s of the hot path on the workloads of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W                       (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W      (the path's CPU implementation on the host cores)
    python bench.py --impl reference-gpu ...                            (the same torch ops on the GPU: cuDNN / cuFFT / ATen)
    python bench.py --config {cfg2,cfg2-B1,cfg1,cfg4-none,cfg4-tv,cfg4-diffusion,cfg5-sweep} ...
    python bench.py --strong --gpus N ...                                (cfg 3: 105 chains fixed, sharded over N ranks)

Default workload (configs[1], "cfg2"): ACDC-shaped 2-D complex MRI, 256x256, 4-coil SENSE, keep-centre column mask
with Bernoulli rate 1/R (R = 40, center_lines_frac = 1/64), NCSNv2Deepest (acdc.yml, ngf 128) with default random
init, step_lr 9e-7, L2Penalty data consistency with lr_scaled 1e6, seg guidance off, 14 chains per GPU.
One "step" = the body of the sampler's inner loop for every chain of the rank (score forward over the real and
imaginary planes of all chains, Langevin update + SENSE proximal step, schedule advance), replayed as one CUDA graph.
`value` = chains * steps / time with the chain state resident in HBM; `e2e` = the same metric through the public
sampler call with the inputs in pinned host memory and the result read back to the host.
Chains shard over ranks with no data-path collective (weak scaling; --strong fixes the total at 105 chains); the only
collective is the posterior mean/std all-reduce after the chains (inside the timed region of --strong).
Every line carries `roofline` (the dominant kernel against the roof that bounds it, SURVEY 8d: convolutions ->
tensor), `cpu_baseline` (the oracle port on the host cores, a bounded sample), and for cfg2 `library_baseline` (the same
torch ops on this GPU) and the SENSE operator GB/s.  The product arm imports the package only; `oracle/` is used
by the reference arms and the cpu_baseline / library_baseline legs alone.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ALD chain-steps/sec"
UNIT = "chain-steps/s"
FLOP_DEEPEST_256 = 838.36e9      # NCSNv2Deepest ngf 128, one 256x256 image; scales with H*W
FLOP_NCSNV2_28 = 15.14e9         # NCSNv2 ngf 128, one 28x28 image
CONFIGS = ("cfg2", "cfg2-B1", "cfg1", "cfg4-none", "cfg4-tv", "cfg4-diffusion", "cfg5-sweep")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-gpu"])
    ap.add_argument("--config", default="cfg2", choices=CONFIGS)
    ap.add_argument("--strong", action="store_true", help="cfg 3: 105 chains in total, sharded over the ranks")
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (cfg2: 14 = 105 chains over 8 GPUs)")
    ap.add_argument("--total-chains", type=int, default=105)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--coils", type=int, default=4)
    ap.add_argument("--R", type=float, default=40.0)
    ap.add_argument("--center-frac", type=float, default=1 / 64)
    ap.add_argument("--e2e-levels", type=int, default=10)
    ap.add_argument("--ref-sample-chains", type=int, default=2, help="chains per step the CPU reference arm actually runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    args = ap.parse_args()
    if args.chains is None:
        args.chains = {"cfg2": 14, "cfg2-B1": 1, "cfg1": 16}.get(args.config, 1)
    return args


# ------------------------------------------------------------------------------------------------ small helpers
def ns(**kw):
    return types.SimpleNamespace(**kw)


def make_config(dataset, ngf, image_size, num_classes, sigma_begin, sigma_end=0.01, device="cpu"):
    """The nested-Namespace config the reference reads (ncsn/configs/*.yml), reduced to the keys the path uses."""
    import torch
    model = ns(sigma_begin=sigma_begin, num_classes=num_classes, sigma_end=sigma_end, sigma_dist="geometric",
               normalization="InstanceNorm++", nonlinearity="elu", ngf=ngf, ema=True, ema_rate=0.999, spec_norm=False)
    data = ns(dataset=dataset, image_size=image_size, channels=1, logit_transform=False, uniform_dequantization=False,
              gaussian_dequantization=False, random_flip=True, rescaled=False)
    recons = ns(sigma_dist="geometric", sigma_begin=sigma_begin, num_classes=num_classes, sigma_end=sigma_end)
    return ns(model=model, data=data, recons=recons, device=torch.device(device))


def phantom(seed, *shape):
    """magnitude U[0,1) with a random phase (stands in for `add_phase`, helpers/load_data.py:372-387)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * torch.exp(1j * torch.randn(*shape, generator=g))).to(torch.complex64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  start() BEFORE the warm-up steps (nvidia-smi needs a few
    hundred ms to produce its first line -- longer than a 10-step timed region when eight ranks start one each), mark() when
    the timed region begins, stop() when it ends: the reported numbers are the samples received inside [mark, stop]; if the
    region was shorter than one period, the samples of the warm-up (the same steps, the same load) are used and the
    window says so."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index, self.t_mark = [], None, index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=3.0):
        """block (bounded) until nvidia-smi has produced its first line, so that the timed region is covered"""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)
        return self

    def mark(self):
        self.t_mark = time.perf_counter()
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        self.proc.terminate()
        t_mark = self.t_mark if self.t_mark is not None else 0.0
        inside = [ln for (t, ln) in self.lines if t_mark <= t <= t_end]
        window = "timed region"
        if not inside:
            inside = [ln for (t, ln) in self.lines if t <= t_end]
            window = "warm-up + timed region (timed region shorter than one sampling period)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in inside:
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
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def matmul_peaks(torch, dev):
    """In-run dense matmul throughput of this GPU for the operand kinds SURVEY 8(d) names: 8192^3, best of 10 (burst) --
    the way MEASURED_PEAKS.json's bf16 figure was taken -- for bf16, fp16 (the kind the convolutions use) and TF32."""
    n = 8192
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    old = torch.backends.cuda.matmul.allow_tf32
    for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("fp16", torch.float16, False), ("tf32", torch.float32, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.randn(n, n, device=dev, dtype=dt)
        b = torch.randn(n, n, device=dev, dtype=dt)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            torch.cuda.synchronize()
            e0.record(); a @ b; e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_tflops_burst"] = 2.0 * n ** 3 / (best / 1e3) / 1e12
        del a, b
    torch.backends.cuda.matmul.allow_tf32 = old
    torch.cuda.empty_cache()
    return out


def conv_traffic_from_profile(N, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from this round's ncu --set full capture
    (profiles/r02_ncu_conv_dominant.json, written by tools/ncu_table.py from the .ncu-rep); None when the capture is of
    another shape or absent -- never a constant typed into this file."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_conv_dominant.json")
    try:
        with open(path) as f:
            rec = json.load(f)
        if rec.get("images") == N and rec.get("size") == n:
            return rec.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ oracle legs (reference arms, cpu_baseline)
def oracle_cfg2_step(args, device, chains, dtype="fp32"):
    """Returns (step_fn, lines): one cfg-2 ALD step of `chains` chains with the oracle's torch ops on `device`."""
    import torch
    from oracle import mri_ops as M, scorenet as SN, ald as OALD
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsnv2 import NCSNv2Deepest
    n = args.size
    cfg = make_config("ACDC", 128, n, 2311, 348.0, device="cpu")
    torch.manual_seed(0)
    net = NCSNv2Deepest(cfg)          # parameter container only (default init); the oracle does the arithmetic
    half = dtype == "fp16"
    Pd = {k: v.detach().to(device) for k, v in net.state_dict().items()}
    if half:
        Pd = {k: (v.half() if v.is_floating_point() and k != "sigmas" else v) for k, v in Pd.items()}
    maps = M.exp_coil_maps(args.coils, n, n, 0).to(device)
    mask = M.keep_center_mask(n, args.R, args.center_frac, seed=0).to(device)
    y = M.sense_forward(phantom(1, 1, 1, n, n).to(device), maps, mask).repeat(1, chains, 1, 1, 1)
    sig = Pd["sigmas"].float()
    x = M.sense_adjoint(y, maps)
    st = {"xr": x.real.float().contiguous(), "xi": x.imag.float().contiguous()}
    labels = torch.zeros(chains, dtype=torch.long, device=device)
    step = 9e-7 * (sig[0] / sig[-1]) ** 2
    Pf = dict(Pd, sigmas=sig)

    def score(v):
        if half:
            return SN.score_forward("NCSNv2Deepest", Pf, v.half().contiguous(memory_format=torch.channels_last), labels).float()
        return SN.score_forward("NCSNv2Deepest", Pf, v, labels)

    def step_fn():
        with torch.no_grad():
            gr = score(st["xr"])
            gi = score(st["xi"])
            xr = OALD.langevin_update(st["xr"], gr, torch.randn_like(st["xr"]), step)
            xi = OALD.langevin_update(st["xi"], gi, torch.randn_like(st["xi"]), step)
            z = M.l2_prox_sense_closed_form(xr + 1j * xi, y, maps, mask, 9e-7 * 1e6, 1.0)
            st["xr"], st["xi"] = z.real.float().contiguous(), z.imag.float().contiguous()
    return step_fn, int(mask.sum())


def oracle_cfg1_step(device):
    import torch
    from oracle import scorenet as SN, ald as OALD
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsnv2 import NCSNv2
    cfg = make_config("MNIST", 128, 28, 232, 50.0, device="cpu")
    torch.manual_seed(0)
    net = NCSNv2(cfg)
    Pd = {k: v.detach().to(device) for k, v in net.state_dict().items()}
    sig = Pd["sigmas"]
    st = {"x": torch.rand(16, 1, 28, 28, device=device)}
    labels = torch.zeros(16, dtype=torch.long, device=device)
    step = 6.2e-6 * (sig[0] / sig[-1]) ** 2

    def step_fn():
        with torch.no_grad():
            g = SN.score_forward("NCSNv2", Pd, st["x"], labels)
            st["x"] = OALD.langevin_update(st["x"], g, torch.randn_like(st["x"]), step)
    return step_fn


def oracle_cfg4_step(device, frames, mode_T, patches):
    """One cfg-4 step on a SAMPLE: `frames` of the 24 frames for the spatial prior + prox, `patches` of the 512
    patches for the learned temporal prior (every frame / patch costs the same: the caller scales)."""
    import torch
    from oracle import mri_ops as M, scorenet as SN, ald as OALD
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsnv2 import NCSNv2Deepest
    n = 128
    cfg = make_config("CINE127", 128, n, 1000, 60.0, device="cpu")
    torch.manual_seed(0)
    net = NCSNv2Deepest(cfg)
    Pd = {k: v.detach().to(device) for k, v in net.state_dict().items()}
    maps = M.exp_coil_maps(4, n, n, 0).to(device)
    mask = M.live_sense_mask(n, 0)[:frames].to(device)                  # (frames,1,1,W)
    y = M.sense_forward(phantom(3, frames, 1, n, n).to(device), maps, mask)
    x = M.sense_adjoint(y, maps)
    st = {"xr": x.real.float().contiguous(), "xi": x.imag.float().contiguous()}
    labels = torch.zeros(frames, dtype=torch.long, device=device)
    sig = Pd["sigmas"]
    step = 1e-4 * (sig[0] / sig[-1]) ** 2
    PT = None
    if mode_T == "diffusion":
        from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
        cfg_T = make_config("CINE127", 128, 24, 400, 40.0, device="cpu")
        cfg_T.data.channels, cfg_T.data.channels_3d = 64, 1
        torch.manual_seed(1)
        PT = {k: v.detach().to(device) for k, v in NCSN3DShallow(cfg_T).state_dict().items()}
        vol = torch.rand(patches, 1, 8, 8, 24, device=device)
        lab_T = torch.zeros(patches, dtype=torch.long, device=device)

    def step_fn():
        with torch.no_grad():
            gr = SN.score_forward("NCSNv2Deepest", Pd, st["xr"], labels)
            gi = SN.score_forward("NCSNv2Deepest", Pd, st["xi"], labels)
            xr = OALD.langevin_update(st["xr"], gr, torch.randn_like(st["xr"]), step)
            xi = OALD.langevin_update(st["xi"], gi, torch.randn_like(st["xi"]), step)
            if mode_T == "tv":
                xr = xr + OALD.temporal_tv_grad(xr.reshape(1, frames, n, n), 0.01).reshape(xr.shape)
                xi = xi + OALD.temporal_tv_grad(xi.reshape(1, frames, n, n), 0.01).reshape(xi.shape)
            if PT is not None:
                SN.score_forward_3d_shallow(PT, vol, lab_T)
            z = M.l2_prox_sense_closed_form(xr + 1j * xi, y, maps, mask, 1e-4, 1.0)
            st["xr"], st["xi"] = z.real.float().contiguous(), z.imag.float().contiguous()
    return step_fn


def time_host(fn, warm, timed):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(timed):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts)


def cpu_leg(args, warm, timed):
    """(chain-steps/s of the oracle port on the host cores, seconds one timed sample step took, description of the bounded
    sample) for args.config."""
    import torch
    torch.set_num_threads(os.cpu_count())
    cores = os.cpu_count()
    cfgname = args.config
    if cfgname in ("cfg2", "cfg2-B1"):
        sample = max(1, min(args.ref_sample_chains, args.chains))
        fn, _ = oracle_cfg2_step(args, "cpu", sample)
        sec = time_host(fn, warm, timed)
        return sample / sec, sec, (f"{timed} timed cfg-2 ALD steps of {sample} of the {args.chains} chains per GPU, batched (every chain costs the same: "
                              f"2 score forwards at {args.size}^2 + update + prox per chain), torch CPU {cores} threads, after {warm} warm-up")
    if cfgname == "cfg1":
        fn = oracle_cfg1_step("cpu")
        sec = time_host(fn, warm, timed)
        return 16 / sec, sec, f"{timed} timed cfg-1 steps of the whole batch (16 images of 28x28), torch CPU {cores} threads, after {warm} warm-up"
    mode = cfgname.split("-")[1]
    frames, patches = 4, 32
    fn = oracle_cfg4_step("cpu", frames, mode, patches)
    fn_sp = oracle_cfg4_step("cpu", frames, "none", patches) if mode == "diffusion" else None
    sec = time_host(fn, warm, timed)
    if fn_sp is not None:      # spatial part scales with frames (x6), temporal prior with patches (x16)
        sec_sp = time_host(fn_sp, warm, timed)
        full = sec_sp * (24 / frames) + max(sec - sec_sp, 0.0) * (512 / patches)
    else:
        full = sec * (24 / frames)
    return 1.0 / full, sec, (f"{timed} timed cfg-4 steps on a sample ({frames} of 24 frames" + (f", {patches} of 512 temporal patches" if mode == "diffusion" else "") +
                        f"), scaled to the whole volume (per-frame / per-patch cost is uniform), torch CPU {cores} threads, after {warm} warm-up")


def cpu_side_legs(args):
    """SURVEY 8(d) CPU legs beside the step: (i) SENSE forward + adjoint of the oracle port at a cfg-5 point, (ii) one
    NCSNv2Deepest forward of 2 images at the benchmark size.  Bounded: a few seconds on 16 cores."""
    import torch
    from oracle import mri_ops as M, scorenet as SN
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsnv2 import NCSNv2Deepest
    torch.set_num_threads(os.cpu_count())
    n, coils, batch = args.size, args.coils, 16
    maps = M.exp_coil_maps(coils, n, n, 0)
    mask = M.keep_center_mask(n, args.R, args.center_frac, seed=0)
    g = torch.Generator().manual_seed(5)
    x = torch.complex(torch.randn(batch, 1, n, n, generator=g), torch.randn(batch, 1, n, n, generator=g))
    S = M.sense_forward(x, maps, mask)
    sec_f = time_host(lambda: M.sense_forward(x, maps, mask), 1, 3)
    sec_a = time_host(lambda: M.sense_adjoint(S, maps), 1, 3)
    alg = 8.0 * batch * n * n * (1 + coils) + 4.0 * coils * n * n
    cfg = make_config("ACDC", 128, n, 2311, 348.0, device="cpu")
    torch.manual_seed(0)
    net = NCSNv2Deepest(cfg)
    Pd = {k: v.detach() for k, v in net.state_dict().items()}
    Pd["sigmas"] = Pd["sigmas"].float()
    xin = torch.rand(2, 1, n, n, generator=g)
    labels = torch.zeros(2, dtype=torch.long)
    with torch.no_grad():
        sec_n = time_host(lambda: SN.score_forward("NCSNv2Deepest", Pd, xin, labels), 1, 2)
    return {"sense_fwd_adj": {"point": f"{coils} coils, {n}x{n}, batch {batch}, R={args.R:g}", "forward_ms": sec_f * 1e3, "adjoint_ms": sec_a * 1e3,
                              "forward_gbs": alg / sec_f / 1e9, "adjoint_gbs": alg / sec_a / 1e9},
            "ncsnv2deepest_forward_2_images": {"size": n, "ms": sec_n * 1e3, "tflops": 2 * 838.36e9 * (n / 256) ** 2 / sec_n / 1e12}}


def cpu_sense_forward(warm, timed):
    """The sweep's headline point on the host cores: two of its 64 images per timed step (per-image cost is uniform)."""
    import torch
    from oracle import mri_ops as M
    torch.set_num_threads(os.cpu_count())
    coils, n, batch, R = 32, 512, 2, 40.0
    maps = M.exp_coil_maps(coils, n, n, 0)
    mask = M.keep_center_mask(n, R, 1 / 64, seed=0)
    g = torch.Generator().manual_seed(5)
    x = torch.complex(torch.randn(batch, 1, n, n, generator=g), torch.randn(batch, 1, n, n, generator=g))
    sec = time_host(lambda: M.sense_forward(x, maps, mask), warm, timed)
    alg = 8.0 * batch * n * n * (1 + coils) + 4.0 * coils * n * n
    sample = (f"{timed} timed SENSE forwards of 2 of the 64 images at 32 coils x 512x512, R=40 (oracle port: torch FFT + coil multiply "
              f"+ mask, {os.cpu_count()} threads), after {warm} warm-up")
    return alg / sec / 1e9, sec, sample


def run_reference(args):
    """`--impl reference`: the path's CPU implementation (the oracle port; the reference itself cannot travel to the box) on the
    host cores with every thread, on the native arm's config / metric / unit; every one of the K timed steps is a bounded
    sample of the workload (ref_sample_chains of the chains, batched)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == "cfg5-sweep":
        val, sec, sample = cpu_sense_forward(args.warmup, args.steps)
        print(json.dumps({"impl": "reference", "metric": "SENSE forward algorithmic GB/s", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, None),
                          "cpu_baseline": {"value": val, "unit": "GB/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    val, sec, sample = cpu_leg(args, args.warmup, args.steps)
    cfg = workload_config(args, None)
    # ms_per_step is what one TIMED step took (a bounded sample of the workload, see cpu_baseline.sample): steps * ms_per_step
    # is this run's timed wall time; `value` is the per-chain rate, which does not depend on how many chains a step holds
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def library_leg(args, torch, dev, chains, kinds=("fp32", "tf32", "fp16")):
    """The same cfg-2 step with the reference's own torch ops on THIS GPU (cuDNN convolutions, cuFFT, ATen elementwise): the
    'library' bar of SURVEY 2.2 / 8(d).  fp32 = cuDNN without TF32, tf32 = TF32 convolutions allowed, fp16 = half weights and
    activations in channels_last (lossy InstanceNorm: a speed bar only)."""
    out = {}
    old_c, old_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.benchmark = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for kind in kinds:
        torch.backends.cudnn.allow_tf32 = kind != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = kind != "fp32"
        try:
            fn, _ = oracle_cfg2_step(args, dev, chains, "fp16" if kind == "fp16" else "fp32")
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out[kind] = {"value": chains / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
            del fn
        except Exception as exc:      # e.g. out of memory at fp32
            out[kind] = {"error": str(exc)[:200]}
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_c, old_m
    out["what"] = (f"oracle port's torch ops on the GPU (cuDNN / cuFFT / ATen, eager, {chains} chains, 3 timed steps after 2): the reference's "
                   "own code path as it would run on this box")
    return out


def run_reference_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config not in ("cfg2", "cfg2-B1"):
        print(json.dumps({"impl": "reference-gpu", "unavailable": "library arm is implemented for cfg2 only"}), flush=True)
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib = library_leg(args, torch, dev, args.chains)
    best = max((v["value"] for v in lib.values() if isinstance(v, dict) and "value" in v), default=None)
    head = lib.get("tf32") or {}
    out = {"impl": "reference-gpu", "metric": METRIC, "value": head.get("value", best), "unit": UNIT, "n_gpus": 1,
           "steps": 3, "warmup": 2, "ms_per_step": head.get("ms_per_step"), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "tf32", "data": "synthetic", "config": workload_config(args, None),
           "e2e": {"value": head.get("value", best), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "note": "state resident on the device (eager torch ops); no host copies in the timed region"},
           "library_baseline": lib, "gpu_launches": None,
           "note": "library arm: none of this repo's kernels; the headline value is the TF32 leg (torch's default conv precision "
                   "is TF32-off = 'fp32' leg; the fp16 channels-last leg is the fastest library setting tried)"}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ workload descriptions
def workload_config(args, line_count):
    c = args.config
    if c in ("cfg2", "cfg2-B1"):
        lines = f"{line_count} of {args.size} lines sampled" if line_count is not None else "keep-centre mask"
        return {"workload": f"{'cfg3 (strong scaling of cfg2)' if args.strong else c}: ACDC-shaped {args.size}x{args.size} complex, {args.coils}-coil SENSE, "
                            f"R={args.R:g} (center_lines_frac={args.center_frac:.5f}: {lines}), NCSNv2Deepest ngf128, ALD + L2Penalty prox",
                "chains_per_gpu": args.chains, "images_per_forward": 2 * args.chains, "n_steps_each": 3, "step_lr": 9e-7, "lr_scaled": 1e6,
                "schedule_levels": 2311, "operand_dtype": "f16 operands and residual stream (fp32 accumulate, fp32 epilogue arithmetic)",
                "l2": "activations (>= 0.45 GB per tensor at 14 chains) exceed the 126 MB L2; no flush needed" if args.chains >= 4 else
                      "2 images per forward: the deep layers are L2-resident, as they are in a real single-chain run",
                "e2e_call_steps": args.e2e_levels * 3}
    if c == "cfg1":
        return {"workload": "cfg1: MNIST-shaped 28x28, NCSNv2 ngf128 unconditional ALD, batch 16 (mnist.yml: L=232, 5 steps each, step_lr 6.2e-6)",
                "chains_per_gpu": 16, "images_per_forward": 16, "n_steps_each": 5, "step_lr": 6.2e-6, "schedule_levels": 232,
                "operand_dtype": "f16 (fp32 accumulate)", "l2": "working set is L2-resident (16 images of 28x28), as in the real run",
                "e2e_call_steps": 50}
    if c.startswith("cfg4"):
        return {"workload": f"{c}: CINE127-shaped 2D+time (1,24,1,128,128), 4-coil SENSE, live 24-frame 'R=16' mask, NCSNv2Deepest ngf128, "
                            f"ALD2DTime mode_T={c.split('-')[1]}" + (" (NCSN3DShallow ngf128 on 512 patches of 8x8x24)" if c.endswith("diffusion") else ""),
                "chains_per_gpu": 1, "images_per_forward": 48, "n_steps_each": 3, "step_lr": 1e-4, "schedule_levels": 1000,
                "operand_dtype": "f16 (fp32 accumulate)", "l2": "activations of 48 frames at 128^2 (0.4 GB per tensor) exceed L2",
                "e2e_call_steps": 30}
    return {"workload": "cfg5: SENSE forward / adjoint / fused-step sweep (coils 4-32, 128^2-512^2, batch 1-64) vs the HBM roofline",
            "l2": "flush between timed iterations: 256 MB written, then 256 MB read (cold and clean L2)"}


# ------------------------------------------------------------------------------------------------ SENSE operator points
def sense_point(torch, P, _lib, L, dev, hbm, coils, size, batch, R, frac):
    """SENSE forward / adjoint / fused ALD step at one cfg-5 sweep point: algorithmic GB/s (fwd/adj
    8N(1+Nc) + 4*Nc*H*W, fused step 32N + 4*Nc*H*W bytes, N = B*H*W) over CUDA-event time, L2 flushed between
    iterations, best of 5.  `adjoint` is A^H on masked data (mask applied); `conj_op_unmasked` is the reference's
    SENSE.conj_op signature, which transforms every column (quirk Q3)."""
    A = P.SENSE("exp", coils, R, frac, (1, size, size), 0)
    A.random_under_fourier.mask = P.keep_center_mask(size, R, frac, seed=0)
    x = torch.randn(batch, 1, size, size, dtype=torch.complex64, device=dev)
    S = A(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)     # 256 MB that is only ever read
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def flush_l2():
        """cold AND clean L2: 256 MB written (everything evicted), then another 256 MB read, so that the write-back of the
        write sweep's dirty lines (~126 MB) is not charged to the timed kernel (2-4 % at the 134 MB point)"""
        flush.zero_()
        flush_rd.sum()
        torch.cuda.synchronize()

    def best(fn):
        """best of 5 CUDA-event times of ONE replay of fn captured in a CUDA graph (the way the samplers run these calls: a
        captured step has no host launch gaps; at the small sweep points two eager launches cost more host time than the
        kernels take), L2 flushed before each"""
        fn(); torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        ts = []
        for _ in range(5):
            flush_l2()
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        del g
        return min(ts)

    N = batch * size * size
    b_fa = 8 * N * (1 + coils) + 4 * coils * size * size
    b_st = 32 * N + 4 * coils * size * size
    state = torch.randn(2, batch, size, size, device=dev); grad = torch.randn_like(state); bvec = torch.randn_like(state)
    mre, _ = A.device_maps(dev)
    plan = A.device_plan(dev, size)
    sc = _lib.AldScalars(0.1, 0.4, 0.01, 1.0)
    step = lambda: _lib.check(L.ipdm_ald_sense_step_plan(plan.handle, state.data_ptr(), grad.data_ptr(), None, bvec.data_ptr(), mre.data_ptr(), None,
                                                         coils, batch, size, sc, None, None, _lib.rng(1, 0), _lib.stream()))
    t = {"forward": best(lambda: A(x)), "adjoint": best(lambda: A.conj_op_masked(S)), "conj_op_unmasked": best(lambda: A.conj_op(S)),
         "fused_ald_step": best(step)}
    out = {"point": f"{coils} coils, {size}x{size}, batch {batch}, R={R:g} ({int(A.random_under_fourier.mask.sum())} lines), k-space {8 * coils * N / 1e6:.0f} MB",
           "coils": coils, "size": size, "batch": batch, "R": R, "lines": int(A.random_under_fourier.mask.sum()), "pruned_kernels": bool(plan.pruned),
           "hbm_peak_gbs": hbm, "l2": "flush between timed iterations: 256 MB written, then 256 MB read (cold and clean L2)", "timing": "CUDA events around one graph replay, best of 5"}
    for k, ms in t.items():
        byt = b_st if k == "fused_ald_step" else b_fa
        out[k] = {"ms": ms, "gbs": byt / ms / 1e6, "frac": byt / ms / 1e6 / hbm}
    del x, S, state, grad, bvec, flush
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ native arm
def build_cfg2(args, P, torch, dev, B):
    n = args.size
    cfg = make_config("ACDC", 128, n, 2311, 348.0, device=str(dev))
    torch.manual_seed(0)
    net = P.NCSNv2Deepest(cfg).to(dev).eval()
    A = P.SENSE("exp", args.coils, args.R, args.center_frac, (1, n, n), 0)
    A.random_under_fourier.mask = P.keep_center_mask(n, args.R, args.center_frac, seed=0)
    y1 = A(phantom(1, 1, 1, n, n).to(dev))                       # (Nc,1,1,H,W)
    meas_host = y1.repeat(1, max(B, 1), 1, 1, 1).cpu().pin_memory()
    sig = P.get_sigmas(cfg, mode="recons")
    params = {"n_steps_each": 3, "step_lr": 9e-7, "denoise": True, "final_only": True}

    def make_sampler(sigmas, measurement, prm=params):
        return P.ALD.ALDInvSegProximalRealImag(P.L2Penalty(A), 1.0, "linear", (B, 1, n, n), net, sigmas, prm, cfg,
                                               measurement=measurement, linear_tfm=A, seg=None, device=dev)
    return ns(n=n, cfg=cfg, net=net, A=A, meas_host=meas_host, sig=sig, params=params, make_sampler=make_sampler,
              lines=int(A.random_under_fourier.mask.sum()), flop_per_chain_step=2 * FLOP_DEEPEST_256 * (n / 256) ** 2)


def time_replays(torch, dist, world, step, warmup, steps, local):
    clocks = ClockSampler(local).start().wait_first()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    clocks.mark()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    info = clocks.stop()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), info


def conv_roofline(args, torch, _lib, L, dev, N, n, ms_step_total, steps, flop_step):
    """The dominant kernel by launch-list share (profiles/): k_conv_halo, 128->128 3x3 at the full image size, in its
    residual variant (RCU / CRP second convolutions) -- timed with CUDA events in this run on tensors of the step's own
    size.  Judged on the TENSOR roof (convolutions are the dense contraction of the path); the HBM view of the
    same launch is the side note."""
    import ctypes
    burst, sustained, hbm, peak_kind = measured_peaks()
    x16 = torch.randn(N, n, n, 128, device=dev).half()
    w16 = (torch.randn(128, 9, 128, device=dev) / 34).half()
    o16 = torch.empty_like(x16)
    o32 = torch.empty(N, n, n, 128, device=dev)
    res = torch.randn(N, n, n, 128, device=dev)
    raw16 = torch.empty_like(x16)
    res16 = res.half()
    conv_flop = 2.0 * N * n * n * 128 * 128 * 9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

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

    # the launch the network issues (16-bit residual stream: f16 in, f16 residual, f16 result, f16 ELU copy = 8 B / element),
    # the same convolution with the fp32 stream of round 1 (12 B / element), and the store-only variant (4 B / element)
    d_t16 = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, None, None, o16.data_ptr(), None, N, n, n, 128, 128, 9, 1, _lib.CONV_F16_ELU,
                          0, 0, res16.data_ptr(), raw16.data_ptr())
    d_res = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, res.data_ptr(), o32.data_ptr(), o16.data_ptr(), None,
                          N, n, n, 128, 128, 9, 1, _lib.CONV_F16_ELU)
    d_f16 = _lib.ConvDesc(x16.data_ptr(), w16.data_ptr(), None, None, None, o16.data_ptr(), None, N, n, n, 128, 128, 9, 1, _lib.CONV_F16_ELU)
    ms_t16, ms_res, ms_f16 = time_conv(d_t16), time_conv(d_res), time_conv(d_f16)
    achieved = conv_flop / (ms_t16 / 1e3) / 1e12
    alg_bytes = N * n * n * 128 * (2 + 2 + 2 + 2)
    step_tflops = flop_step / (ms_step_total / steps / 1e3) / 1e12
    view = lambda ms: {"ms_per_launch": ms, "achieved": conv_flop / (ms / 1e3) / 1e12, "peak": burst, "unit": "TFLOP/s",
                       "frac": conv_flop / (ms / 1e3) / 1e12 / burst, "bound": "tensor"}
    out = {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
           "kernel": "k_conv_halo<residual f16 + result f16 + ELU copy f16> (128->128 3x3 @%dx%d, %d images)" % (n, n, N),
           "traffic": conv_traffic_from_profile(N, n),
           "ms_per_launch": ms_t16, "flop_per_launch": conv_flop,
           "peak_source": f"MEASURED_PEAKS.json bf16 burst TFLOP/s ({peak_kind}); the in-run fp16 figure is under `matmul_peaks`",
           "hbm_view": {"algorithmic_bytes_per_launch": alg_bytes, "gbs": alg_bytes / ms_t16 / 1e6, "frac_of_copy_peak": alg_bytes / ms_t16 / 1e6 / hbm,
                        "arithmetic_intensity_flop_per_byte": conv_flop / alg_bytes, "ridge_flop_per_byte": burst * 1e12 / (hbm * 1e9)},
           "same_shape_fp32_residual_stream": view(ms_res),
           "same_shape_f16_store_only": view(ms_f16),
           "whole_step_conv_tflops": step_tflops, "whole_step_frac_of_sustained": step_tflops / sustained,
           "whole_step_note": ("algorithmic FLOPs of the reference's layers (SURVEY A.1: 838.36 GFLOP per 256x256 forward) over the step time.  "
                               "The kernels EXECUTE 0.952 of them: the three ConvMeanPool layers run as 4x4 stride-2 convolutions (16/36 of "
                               "their tap evaluations) and the three pooled 1x1 shortcuts on pooled operands (1/4), DESIGN 4.1"),
           "executed_flop_fraction": 0.952,
           "whole_step_frac_of_burst": step_tflops / burst}
    del raw16, res16
    del x16, w16, o16, o32, res
    torch.cuda.empty_cache()
    return out


def run_native(args):
    import torch
    import torch.distributed as dist
    from inverseproblemwithdiffusionmodel_b200 import _lib, chains as CH
    from inverseproblemwithdiffusionmodel_b200.ncsn.linear_transforms.undersampling_fourier import SENSE, keep_center_mask
    from inverseproblemwithdiffusionmodel_b200.ncsn.models import get_sigmas
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsnv2 import NCSNv2, NCSNv2Deepest
    from inverseproblemwithdiffusionmodel_b200.ncsn.models.proximal_op import L2Penalty
    from inverseproblemwithdiffusionmodel_b200.ncsn.models import ALD_optimizers as ALD
    P = ns(SENSE=SENSE, keep_center_mask=keep_center_mask, get_sigmas=get_sigmas, NCSNv2=NCSNv2, NCSNv2Deepest=NCSNv2Deepest,
           L2Penalty=L2Penalty, ALD=ALD)
    rank, local, world = CH.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    burst, sustained, hbm, _ = measured_peaks()
    if args.config == "cfg5-sweep":
        return run_sweep(args, torch, P, _lib, L, dev, hbm, rank, world, dist)
    if args.config in ("cfg2", "cfg2-B1"):
        return run_cfg2(args, torch, dist, P, _lib, CH, L, dev, rank, local, world)
    return run_other(args, torch, dist, P, _lib, CH, L, dev, rank, local, world)


def run_cfg2(args, torch, dist, P, _lib, CH, L, dev, rank, local, world):
    import numpy as np
    burst, sustained, hbm, _ = measured_peaks()
    n = args.size
    seed = 1234                                                    # one seed on every rank: chains differ by their global id
    if args.strong:
        mine = CH.chain_partition(args.total_chains, world, rank)   # round-robin: 105 -> 14,13,...,13 on 8 ranks
        total = args.total_chains
    else:
        mine = CH.chain_partition(args.chains * world, world, rank)
        total = args.chains * world
    B = len(mine)
    if args.strong:
        args.chains = -(-total // world)
    W = build_cfg2(args, P, torch, dev, max(B, 1))
    kw = dict(label=None, lamda=1.0, save_dir="/tmp", lr_scaled=1e6, seg_mode="full", seed=seed, chain_ids=mine if B else [0])

    # ---- device-resident throughput: K replays of the captured step -------------------------------
    sampler = W.make_sampler(W.sig, W.meas_host.to(dev))
    chain = sampler(return_chain=True, **kw)
    step = chain["step"]
    launches_per_step = step.launches
    run_step = step if B > 0 else (lambda: None)                   # a rank without chains idles but joins every collective
    stats = CH.PosteriorStats(n * n, dev)

    def posterior_tail(st=None):
        """what cfg 3 does after the chains: statistics of this rank's chains, ONE all-reduce, per-pixel mean / std"""
        st = stats if st is None else st
        if B > 0:
            st.add(torch.complex(chain["state"][0], chain["state"][1]).reshape(B, 1, n, n))
        st.all_reduce()
        return st.finalize((n, n))

    if args.strong:
        # the collective and the reduction are INSIDE the timed region: K steps of every chain + posterior statistics
        clocks = ClockSampler(local).start().wait_first()
        for _ in range(args.warmup):
            run_step()
        posterior_tail(CH.PosteriorStats(n * n, dev))              # warm-up of the tail too (first all-reduce of this size: 90 ms of NCCL set-up)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        clocks.mark()
        e0.record()
        for _ in range(args.steps):
            run_step()
        post = posterior_tail()
        e1.record()
        torch.cuda.synchronize()
        clock_info = clocks.stop()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    else:
        ms_total, clock_info = time_replays(torch, dist, world, run_step, args.warmup, args.steps, local)
        post = posterior_tail()
    value = total * args.steps / (ms_total / 1e3)
    finite = bool(torch.isfinite(chain["state"]).all())

    # ---- end to end through the public call: host measurement in, host result out ------------------
    e2e_sig = torch.tensor(np.exp(np.linspace(np.log(348.0), np.log(0.01), args.e2e_levels))).float().to(dev)
    s2 = W.make_sampler(e2e_sig, W.meas_host, dict(W.params, denoise=False))
    res = s2(**kw)                                                 # first call captures the graph
    torch.cuda.synchronize()
    n_calls = max(1, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n_calls):
        if B > 0:
            res = s2(**kw)
    e1.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    call_steps = args.e2e_levels * 3
    e2e_value = total * call_steps * n_calls / (float(t2.item()) / 1e3)
    h2d = W.meas_host.numel() * 8
    d2h = res[0].numel() * 8
    torch.set_grad_enabled(True)

    roofline = conv_roofline(args, torch, _lib, L, dev, 2 * max(B, 1), n, ms_total, args.steps, max(B, 1) * W.flop_per_chain_step)

    # ---- SENSE operator GB/s (second headline metric) at cfg-5 points whose k-space exceeds L2 ------------
    sense, lib_base, peaks = None, None, None
    if rank == 0:
        del sampler, s2
        torch.cuda.empty_cache()
        sense = sense_point(torch, P, _lib, L, dev, hbm, coils=args.coils, size=n, batch=64, R=args.R, frac=args.center_frac)
        if world == 1:   # the largest cfg-5 sweep point (k-space 4.3 GB >> L2): where SURVEY 8(d) evaluates the HBM fraction
            sense["largest_sweep_point"] = sense_point(torch, P, _lib, L, dev, hbm, coils=32, size=512, batch=64, R=40.0, frac=1 / 64)
            peaks = matmul_peaks(torch, dev)
            if not args.no_library_baseline:
                lib_base = library_leg(args, torch, dev, max(B, 1), kinds=("tf32", "fp16"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, _, sample = cpu_leg(args, 1, 2)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
        try:
            cpu["legs"] = cpu_side_legs(args)
        except Exception as e:      # the legs are side information: never lose the bench line over them
            cpu["legs"] = {"error": repr(e)[:200]}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
           "dtype": "f16", "data": "synthetic", "config": workload_config(args, W.lines),
           "clocks": clock_info,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "note": f"one bench step = one public sampler call of {call_steps} ALD steps; {n_calls} calls timed"},
           "gpu_launches": launches_per_step * args.steps,
           "launches_per_step": launches_per_step,
           "roofline": roofline, "sense": sense, "cpu_baseline": cpu, "library_baseline": lib_base, "matmul_peaks": peaks,
           "state_finite": finite, "posterior_chains": post["n"], "my_chains": B, "total_chains": total}
    if lib_base and "tf32" in lib_base and "value" in lib_base["tf32"]:
        out["vs_library"] = {"tf32": value / lib_base["tf32"]["value"],
                             "fp16_channels_last": value / lib_base["fp16"]["value"] if "value" in lib_base.get("fp16", {}) else None}
    if args.strong:
        out["strong_note"] = ("timed region = K steps of every chain + per-rank posterior statistics + the all-reduce; "
                              f"{total} chains over {world} ranks (round-robin: at most {-(-total // world)} per rank)")
    print(json.dumps(out), flush=True)


def run_other(args, torch, dist, P, _lib, CH, L, dev, rank, local, world):
    """cfg 1 and cfg 4 in the same schema.  Every rank runs the same workload (replicas: weak scaling)."""
    burst, sustained, hbm, _ = measured_peaks()
    c = args.config
    torch.manual_seed(0)
    if c == "cfg1":
        cfg = make_config("MNIST", 128, 28, 232, 50.0, device=str(dev))
        net = P.NCSNv2(cfg).to(dev).eval()
        sig = P.get_sigmas(cfg)
        params = {"n_steps_each": 5, "step_lr": 6.2e-6, "denoise": True, "final_only": True}
        x0_host = torch.rand(16, 1, 28, 28).pin_memory()
        mk = lambda s: P.ALD.ALDUnconditionalSampler((16, 1, 28, 28), net, s, params, cfg, device=dev)
        call = lambda smp: smp(seed=1, x_init=x0_host)
        chains, flop_step = 16, 16 * FLOP_NCSNV2_28
        e2e_sig = sig[::len(sig) // 10][:10]
        h2d = x0_host.numel() * 4
    else:
        n = 128
        mode = c.split("-")[1]
        cfg = make_config("CINE127", 128, n, 1000, 60.0, device=str(dev))
        net = P.NCSNv2Deepest(cfg).to(dev).eval()
        sig = P.get_sigmas(cfg, mode="recons")
        A = P.SENSE("exp", 4, 16, 1 / 8, (1, n, n), 0)              # live mask: 24 frames, the reference's "R = 16" parameters
        meas_host = A(phantom(3, 24, 1, n, n).to(dev)).reshape(4, 1, 24, 1, n, n).cpu().pin_memory()
        params = {"n_steps_each": 3, "step_lr": 1e-4}
        net_T, sig_T, flop_T = None, sig[-10:], 0.0
        if mode == "diffusion":
            from inverseproblemwithdiffusionmodel_b200.ncsn.models.ncsn3d import NCSN3DShallow
            cfg_T = make_config("CINE127", 128, 24, 400, 40.0, device=str(dev))
            cfg_T.data.channels, cfg_T.data.channels_3d = 64, 1
            torch.manual_seed(1)
            net_T = NCSN3DShallow(cfg_T).to(dev).eval()
            sig_T = P.get_sigmas(cfg_T)
            # 3x3x3 convolutions of NCSN3DShallow: 19 x 128->128, 2 x 128->256, 2 x 256->256 at T; 27 x 256->256 at T/2
            flop_T = 512 * 8 * 8 * 2 * 27 * (24 * (19 * 128 * 128 + 2 * 128 * 256 + 2 * 256 * 256) + 12 * 27 * 256 * 256)
        mode_T = {"none": "none", "tv": "tv", "diffusion": "diffusion1d"}[mode]
        mk = lambda s: P.ALD.ALD2DTime(P.L2Penalty(A), net_T, sig_T, (1, 24, 1, n, n), net, s, params, cfg, measurement=meas_host,
                                       linear_tfm=A, device=dev)
        call = lambda smp: smp(save_dir="/tmp", lr_scaled=1.0, mode_T=mode_T, lamda_T=1.0 if mode == "diffusion" else 0.01, seed=2)
        chains, flop_step = 1, 48 * FLOP_DEEPEST_256 * (n / 256) ** 2 + flop_T
        e2e_sig = sig[-10:]                                          # the tail of the schedule: every level runs the temporal step
        h2d = meas_host.numel() * 8
    smp = mk(e2e_sig)
    res = call(smp)                                                  # captures the step graph(s)
    fc = list(smp._fast_cache.values())[0]
    step = fc["step_T"] if "step_T" in fc else fc["step"]
    ms_total, clock_info = time_replays(torch, dist, world, step, args.warmup, args.steps, local)
    value = chains * world * args.steps / (ms_total / 1e3)
    # ---- end to end: the public call on a 10-level schedule, pinned-host inputs in, host result out
    n_calls = max(1, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(n_calls):
        res = call(smp)
    e1.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    call_steps = len(e2e_sig) * params["n_steps_each"]
    e2e_value = chains * world * call_steps * n_calls / (float(t2.item()) / 1e3)
    torch.set_grad_enabled(True)
    step_tflops = flop_step / (ms_total / args.steps / 1e3) / 1e12
    roofline = {"bound": "tensor", "achieved": step_tflops, "peak": sustained, "unit": "TFLOP/s", "frac": step_tflops / sustained,
                "kernel": "all convolution launches of one step (k_conv_halo / k_conv_igemm), algorithmic FLOPs over the step time", "traffic": None,
                "peak_source": "MEASURED_PEAKS.json bf16 sustained TFLOP/s (kernels timed inside a long step)", "flop_per_step": flop_step}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, _, sample = cpu_leg(args, 1, 2)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
    r0 = res[0]
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
           "data": "synthetic", "config": workload_config(args, None), "clocks": clock_info,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": r0.numel() * r0.element_size(),
                   "note": f"one bench step = one public sampler call of {call_steps} ALD steps; {n_calls} calls timed"},
           "gpu_launches": smp.launches_per_step * args.steps, "launches_per_step": smp.launches_per_step,
           "roofline": roofline, "cpu_baseline": cpu,
           "state_finite": bool(torch.isfinite(torch.view_as_real(r0) if r0.is_complex() else r0).all())}
    print(json.dumps(out), flush=True)


def run_sweep(args, torch, P, _lib, L, dev, hbm, rank, world, dist):
    """cfg 5: the SENSE operator sweep.  metric = forward GB/s at the largest point; every point is in `sweep`."""
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cases = [(4, 256, 1, 40), (4, 256, 14, 40), (4, 256, 64, 40), (8, 256, 16, 16), (16, 256, 64, 16), (32, 128, 64, 8), (4, 512, 16, 40),
             (16, 512, 16, 40), (32, 512, 64, 40), (32, 512, 16, 4), (4, 128, 64, 4)]
    before = L.ipdm_launch_count()
    pts = [sense_point(torch, P, _lib, L, dev, hbm, coils=c, size=s, batch=b, R=float(R), frac=1 / 64) for (c, s, b, R) in cases]
    big = [p for p in pts if p["coils"] == 32 and p["size"] == 512 and p["batch"] == 64][0]
    out = {"metric": "SENSE forward algorithmic GB/s", "value": big["forward"]["gbs"], "unit": "GB/s", "n_gpus": 1, "steps": 5, "warmup": 1,
           "ms_per_step": big["forward"]["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": workload_config(args, None),
           "roofline": {"bound": "hbm", "achieved": big["forward"]["gbs"], "peak": hbm, "unit": "GB/s", "frac": big["forward"]["frac"], "traffic": None,
                        "kernel": "kp_fwd_rows + kp_fwd_cols at " + big["point"]},
           "e2e": None, "gpu_launches": int(L.ipdm_launch_count() - before), "sweep": pts}
    if not args.no_cpu_baseline:
        v, _, sample = cpu_sense_forward(1, 3)
        out["cpu_baseline"] = {"value": v, "unit": "GB/s", "cores": os.cpu_count(), "kind": "port", "sample": sample}
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
