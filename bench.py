#!/usr/bin/env python
"""Headline benchmark: Libero goal-video synthesis, 128x128, 7 generated + 1 conditioning frame,
100 denoise steps (the shipped config runs DDPM ancestral sampling), batch 16 per GPU.

  python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --steps K --warmup W  # the reference's own CPU path on the host cores
  python bench.py --impl reference --reference-device cuda --steps 3   # the reference's PyTorch-eager GPU path
                                                         # (E32 / E16); our arm runs this itself
                                                         # in a subprocess at N = 1 (`gpu_eager` block)

One "step" = one GoalGaussianDiffusion.sample() call (100 UNet forwards + sampler updates) on one batch of
synthetic prompts.  Prints ONE JSON line (contract in the task statement / DESIGN.md §6); the last key,
`summary`, repeats the numbers that matter in < 1 kB so a truncated tail still carries them.

The reference arms drive the reference's UNMODIFIED modules when a copy is present (`baseline/_ref`, shipped by
`oracle/make_ref.py`; `/root/reference` in the build container) and fall back to the oracle port otherwise;
`cpu_baseline.kind` / `gpu_eager.kind` say which.  Nothing under oracle/ is touched by the measured path of
our arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "video frames/s (128x128, 7 generated + 1 cond frame, 100 denoise steps)"
UNIT = "frames/s"
FRAMES = 7
H = W = 128
DENOISE_STEPS = 100
TOKENS = 12
FLOP_PER_VIDEO_STEP = 2132.6e9  # algorithmic GFLOP per (video, denoise step)
POLICY_B, POLICY_T, POLICY_DA = 256, 16, 7
POLICY_METRIC = "policy samples/s (train step: fwd + bwd + grad all-reduce + clip + AdamW + EMA)"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(n_gpus, batch):
    return {"workload": "Libero goal-video synthesis 128x128x8 (7 generated + 1 cond), 100 denoise steps "
                        "(DDPM ancestral = the shipped config), Unet_Libero 201M params",
            "global_batch": batch * n_gpus, "batch_per_gpu": batch, "frames": FRAMES, "denoise_steps": DENOISE_STEPS,
            "parallelism": f"dp{n_gpus}", "precision": "bf16x3 split product, fp32 accumulate (fp32-class, <=1e-3)",
            "l2": "working set per denoise step (>1.8 GB/sample) exceeds the 126 MB L2; no explicit flush"}


def perturb_(module, seed: int):
    """Seeded random-init weights (no checkpoints offline): the module's own initialisation + 0.02 N(0, 1) on every
    tensor, because the temporal convs are dirac-initialised and the norms start at gain 1."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.numel():
                p.add_(0.02 * torch.randn(p.shape, generator=g).to(p.device))
    return module


# ---------------------------------------------------------------------------------------------------------------
# reference arms: the reference's own modules (baseline/_ref) on the host cores / on cuda:0 through stock PyTorch
# ---------------------------------------------------------------------------------------------------------------
def _reference_modules():
    """(kind, video_diffusion_factory, policy_factory): the UNMODIFIED reference classes when a copy of the
    reference is present, else None factories (callers fall back to the oracle port)."""
    from oracle import ref_import as R
    if R.available():
        return "reference", R.build_reference_video_diffusion, R.build_reference_policy
    return "port", None, None


def _port_video_step(dev):
    """Oracle-port stand-in for one denoise step at B (only when no copy of the reference is present)."""
    from oracle import video_oracle as VO
    with open(os.path.join(ROOT, "tests", "golden", "goal_diffusion_state_dict_layout.json")) as f:
        lay = json.load(f)
    shapes = {k[len("model."):]: tuple(v) for k, v in lay.items() if k.startswith("model.")}
    sd = {k: v.to(dev) for k, v in VO.seeded_state_dict(shapes, 2).items()}

    def step(x_cond, te):
        B = x_cond.shape[0]
        x = torch.randn(B, 3 * FRAMES, H, W, device=dev)
        t = torch.full((B,), 50, dtype=torch.long, device=dev)
        with torch.no_grad():
            VO.unet_libero_forward(sd, torch.cat([x, x_cond], 1), t, te)
    return step


def _video_step_fn(dev, n_denoise=1):
    """callable(x_cond, te): `n_denoise` denoise steps of the reference's video path at batch B through its own
    public call (`GoalGaussianDiffusion.sample` with a DDIM plan of `n_denoise` steps)."""
    kind, make_video, _ = _reference_modules()
    if make_video is None:
        port = _port_video_step(dev)
        return kind, lambda xc, te: [port(xc, te) for _ in range(n_denoise)]
    diff = perturb_(make_video(timesteps=DENOISE_STEPS, sampling_timesteps=n_denoise), 2).to(dev).eval()
    assert diff.is_ddim_sampling

    def step(x_cond, te):
        with torch.no_grad():
            diff.sample(x_cond, te, batch_size=x_cond.shape[0])
    return kind, step


def _policy_step_fn(dev, B):
    """callable(): one optimisation step of the reference's policy path at batch B: compute_loss + backward +
    clip_grad_norm_(1.0) + AdamW(lr 1e-4, betas (.95, .999), wd 1e-6) (lb_online_trainer_v7.py:598-618)."""
    kind, _, make_policy = _reference_modules()
    g = torch.Generator().manual_seed(5)
    if make_policy is None:
        from oracle import encoder_oracle as EO
        from oracle import policy_oracle as PO
        from v2a_b200.diffusion_policy import build_libero_policy          # SpatialSoftmax buffer constants only
        with open(os.path.join(ROOT, "tests", "golden", "policy_loss_golden_meta.json")) as f:
            layout = json.load(f)["layout"]
        full = dict(build_libero_policy().state_dict())
        full.update(PO.seeded_full_policy_state_dict(layout, 12))
        psd = {}
        for k, v in full.items():
            v = v.detach().to(dev)
            trainable = v.is_floating_point() and v.numel() > 0 and k.rsplit(".", 1)[-1] in ("weight", "bias")
            psd[k] = v.requires_grad_(True) if trainable else v
        usd = {k[len("model."):]: v for k, v in psd.items() if k.startswith("model.")}
        params = [v for v in psd.values() if v.requires_grad]
        obs = {"img_obs_1": torch.rand(B, 3, 128, 128, generator=g).to(dev),
               "img_goal_1": torch.rand(B, 3, 128, 128, generator=g).to(dev)}
        traj = (torch.rand(B, POLICY_T, POLICY_DA, generator=g) * 2 - 1).to(dev)
        acp = PO.ddpm_alphas_cumprod(100).to(dev)

        def loss_fn():
            noise = torch.randn(traj.shape, device=dev)
            tt = torch.randint(0, 100, (B,), device=dev)
            feat = EO.obs_encoder_forward(psd, "obs_encoder.", {k: v * 2 - 1 for k, v in obs.items()})
            return PO.epsilon_loss(usd, traj, feat.float(), noise, tt, acp)
    else:
        policy = perturb_(make_policy(), 12).to(dev).train()
        params = [p for p in policy.parameters() if p.requires_grad and p.numel()]
        batch = {"obs": {"img_obs_1": torch.rand(B, 1, 3, 128, 128, generator=g).to(dev),
                         "img_goal_1": torch.rand(B, 1, 3, 128, 128, generator=g).to(dev)},
                 "action": (torch.rand(B, POLICY_T, POLICY_DA, generator=g) * 2 - 1).to(dev)}
        loss_fn = lambda: policy.compute_loss(batch)
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.95, 0.999), eps=1e-8, weight_decay=1e-6)

    def step():
        opt.zero_grad(set_to_none=True)
        loss_fn().backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
    return kind, step


def run_reference(args):
    """`--impl reference`: the reference's CPU path on all host cores.  Each of the W + K steps is a bounded sample of
    the workload -- ONE denoise step (UNet forward + sampler update through the reference's own `sample()`) of ONE
    video; the metric is extrapolated linearly (x 100 denoise steps; the CPU path gains nothing from batching)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    kind, step = _video_step_fn("cpu")
    g = torch.Generator().manual_seed(0)
    x_cond, te = torch.rand(1, 3, H, W, generator=g), torch.randn(1, TOKENS, 512, generator=g)
    for _ in range(args.warmup):
        step(x_cond, te)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(x_cond, te)
    step_s = (time.perf_counter() - t0) / max(1, args.steps)
    fps = FRAMES / (step_s * DENOISE_STEPS)
    sample = (f"{args.steps} timed steps after {args.warmup} warm-ups; one step = ONE denoise step of ONE video at 128x128x7 "
              f"through the reference's GoalGaussianDiffusion.sample (1-step DDIM plan); a workload step is 100 of these x 16 "
              f"videos: linear extrapolation")
    # the policy path beside it: 1 warm-up + 2 timed optimisation steps at B = 256
    pol = None
    try:
        pk, pstep = _policy_step_fn("cpu", POLICY_B)
        pstep()
        t0 = time.perf_counter()
        for _ in range(2):
            pstep()
        ps = (time.perf_counter() - t0) / 2
        pol = {"metric": POLICY_METRIC, "value": POLICY_B / ps, "unit": "samples/s", "ms_per_step": ps * 1e3,
               "kind": pk, "cores": threads, "sample": "2 timed compute_loss + backward + clip + AdamW steps at B=256"}
    except Exception as exc:       # the policy leg must not cost the line
        pol = {"error": f"{type(exc).__name__}: {exc}"}
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus, 16),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "policy": pol}
    print(json.dumps(line))


def run_reference_gpu_eager(args):
    """`--impl reference --reference-device cuda`: the reference's PyTorch-eager GPU path on cuda:0 -- its own
    modules (baseline/_ref) through stock PyTorch / cuDNN; none of this repo's kernels run here.

    E16: as shipped -- cudnn.benchmark on, TF32 on, torch.autocast(float16) (scripts/train_libero_dp.py:10,25-26).
    E32: fp32, TF32 off, no autocast -- the numerics the 1e-3 parity bar is defined against.
    cudnn.benchmark autotunes during the warm-up steps, outside the timed region (V2A_EAGER_CUDNN_BENCHMARK=0 =
    heuristics, what the parent falls back to when autotuning does not finish inside its time limit)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dry = bool(os.environ.get("V2A_EAGER_DRY_RUN"))    # CPU plumbing check of this function (build container)
    dev = "cpu" if dry else "cuda"
    bench_on = os.environ.get("V2A_EAGER_CUDNN_BENCHMARK", "1") != "0"
    torch.backends.cudnn.benchmark = bench_on
    B = 1 if dry else args.batch
    k = max(1 if dry else 3, min(args.steps, 5))
    g = torch.Generator().manual_seed(0)
    x_cond = torch.rand(B, 3, H, W, generator=g).to(dev)
    te = torch.randn(B, TOKENS, 512, generator=g).to(dev)
    ND = 1 if dry else 2            # denoise steps per sample() call

    def timed(fn, n, warm):
        if dry:
            t0 = time.perf_counter()
            fn()
            return (time.perf_counter() - t0) * 1e3
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def setting(name):
        tf32 = name == "E16"
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        return torch.autocast(dev, dtype=torch.bfloat16 if dry else torch.float16, enabled=(name == "E16"))

    kind, vstep = _video_step_fn(dev, ND)
    video = {}
    for name in ("E16", "E32"):                      # the fast setting first: partial results survive a timeout
        with setting(name):
            ms = timed(lambda: vstep(x_cond, te), k, 2) / ND
        video[name] = {"ms_per_denoise_step": ms, "frames_per_s": B * FRAMES / (ms * 1e-3 * DENOISE_STEPS),
                       "tflops_algorithmic": B * FLOP_PER_VIDEO_STEP / (ms * 1e-3) / 1e12, "timed_steps": k}
        print(json.dumps({"partial": "video", name: video[name]}), file=sys.stderr, flush=True)
    del vstep
    if not dry:
        torch.cuda.empty_cache()
    PB = 2 if dry else POLICY_B
    pkind, pstep = _policy_step_fn(dev, PB)
    policy = {}
    for name in ("E16", "E32"):
        with setting(name):
            ms = timed(pstep, max(1 if dry else 3, k), 2)
        policy[name] = {"ms_per_step": ms, "samples_per_s": PB / (ms * 1e-3)}
        print(json.dumps({"partial": "policy", name: policy[name]}), file=sys.stderr, flush=True)
    line = {"impl": "reference", "device": "cuda", "metric": METRIC, "unit": UNIT, "n_gpus": 1,
            "value": video["E32"]["frames_per_s"], "steps": k, "warmup": 2, "higher_is_better": True, "dtype": "f32",
            "data": "synthetic", "config": workload_config(1, B),
            "gpu_eager": {"kind": kind, "cudnn_benchmark": bench_on, "batch": B, "video": video, "policy": policy,
                          "what": "reference modules on stock PyTorch/cuDNN, cuda:0. video = its sample() with a 2-step "
                                  "DDIM plan at B, per denoise step, x100; policy = compute_loss + backward + clip + "
                                  "AdamW at B=256. E16 = TF32 + fp16 autocast (as shipped), E32 = fp32, TF32 off"},
            "torch": torch.__version__, "gpu": "dry run on cpu" if dry else torch.cuda.get_device_name(0)}
    print(json.dumps(line))


def gpu_eager_subprocess(batch: int, budget_s: float):
    """Run the GPU-eager reference arm in a child process (its cuDNN autotuning cannot be bounded from inside):
    first with cudnn.benchmark on (the shipped setting), then -- if that did not finish -- with heuristics."""
    out = {}
    for bench_on, limit in (("1", budget_s), ("0", min(120.0, budget_s))):
        env = dict(os.environ, V2A_EAGER_CUDNN_BENCHMARK=bench_on)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                "--reference-device", "cuda", "--steps", "3", "--batch", str(batch)],
                               env=env, capture_output=True, text=True, timeout=limit)
            lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
            if p.returncode == 0 and lines:
                out = json.loads(lines[-1])["gpu_eager"]
                out["wall_s"] = round(time.time() - t0, 1)
                return out
            out = {"error": f"rc {p.returncode}: {p.stderr[-300:]}"}
        except subprocess.TimeoutExpired as exc:
            part = [ln for ln in ((exc.stderr or b"").decode(errors="ignore") if isinstance(exc.stderr, bytes)
                                  else (exc.stderr or "")).splitlines() if ln.startswith('{"partial"')]
            out = {"error": f"timed out after {limit:.0f} s with cudnn.benchmark={bench_on}", "partial": part[-4:]}
    return out


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def measure_kernel_roofline(diff, batch, peaks):
    """Per-launch CUDA-event timing of the dominant kernel (igemm) over one eager denoise step."""
    eng = diff.model.unet.engine(batch, FRAMES, H, W, "cuda")
    st = eng._sampler
    eng.bind_static(st["x"], st["cond"], st["v"])
    evs = []
    from v2a_b200 import ops
    classes = [ops.Igemm, ops.IgemmDual]       # IgemmDual = a Conv3d's spatial + temporal GEMM in one launch
    orig = {c: c.run for c in classes}

    def wrap(cls):
        def timed_run(self):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig[cls](self)
            e1.record()
            evs.append((self, e0, e1))
        return timed_run
    for c in classes:
        c.run = wrap(c)
    try:
        for _ in range(3):
            evs.clear()
            eng.run_static()
            torch.cuda.synchronize()
    finally:
        for c in classes:
            c.run = orig[c]
    tot_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in evs)
    executed = sum(g.flops for g, _, _ in evs)          # what the tensor cores did (sub-pixel upsample: 4 of 9 taps)
    algorithmic = sum(g.algo_flops for g, _, _ in evs)  # what the reference's algorithm asks of these launches
    n = len(evs)
    achieved = algorithmic / (tot_ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    traffic, traffic_src = None, None
    for name in ("r2_igemm_traffic.json", "r1_igemm_traffic.json"):   # ncu capture of the same workload (not measurable live)
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tj = json.load(f)
            if batch == 16:
                traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/" + name
            break
        except (OSError, KeyError, ValueError):
            continue
    return {"bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit-GEMM conv)", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "launches_per_denoise_step": n, "avg_launch_ms": tot_ms / n,
            "algorithmic_flop_per_launch": algorithmic / n, "executed_flop_per_launch": executed / n,
            "mma_passes": eng.passes, "tensor_pipe_frac": executed * eng.passes / (tot_ms * 1e-3) / 1e12 / peak,
            "kernel_ms_per_denoise_step": tot_ms}


def run_policy(args, world, rank, barrier, max_over_ranks, peaks):
    """Policy samples/s at B=256 per GPU (weak scaling; the gradient all-reduce of every slab starts from inside
    backward for N>1)."""
    import torch.nn.functional as F
    from v2a_b200 import obs_encoder as OE
    from v2a_b200 import policy_unet1d as PU
    from v2a_b200.diffusion_policy import build_libero_policy
    from v2a_b200.train_step import PolicyTrainStep

    B, T, Da = POLICY_B, POLICY_T, POLICY_DA
    steps, warm = args.policy_steps, max(3, args.warmup)
    torch.manual_seed(77)                                   # same initial weights on every rank (DDP semantics)
    policy = build_libero_policy().to("cuda")
    policy.train()
    g = torch.Generator().manual_seed(2000 + rank)
    host = {"img_obs_1": torch.rand(B, 1, 3, 128, 128, generator=g).pin_memory(),
            "img_goal_1": torch.rand(B, 1, 3, 128, 128, generator=g).pin_memory(),
            "action": (torch.rand(B, T, Da, generator=g) * 2 - 1).pin_memory()}
    loss_host = torch.zeros(1).pin_memory()
    dev = {k: v.cuda() for k, v in host.items()}

    def timed(fn, n):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n

    # (a) the path north_star names: ConditionalUnet1D fwd + bwd (+ all-reduce) + fused optimiser tail
    net = policy.model
    step_u = PolicyTrainStep(net)
    noisy = torch.randn(B, T, Da, device="cuda")
    noise = torch.randn(B, T, Da, device="cuda")
    tt = torch.randint(0, 100, (B,), device="cuda")
    gc = torch.randn(B, 128, device="cuda")
    loss_u = lambda: F.mse_loss(net(noisy, tt, global_cond=gc), noise)
    ms_u = timed(lambda: step_u.step(loss_u), steps)
    eng = PU.last_engine(net)
    flops = sum(gm.flops for gm in eng.igemms) + sum(gm.flops for gm in eng.wgrads)   # forward + dgrad + wgrad GEMMs
    # the forward / backward lists replay as CUDA graphs: count their kernels from the plan
    launches_u = len(eng.fwd) + len(eng.bwd) + len(eng._wchunks) + len(eng._vchunks) + 2
    step_u.close()
    del step_u

    # (b) e2e through the public API: host batch -> compute_loss -> backward -> optimiser -> loss to host
    step_p = PolicyTrainStep(policy)

    def e2e_step():
        b = {"obs": {"img_obs_1": host["img_obs_1"].to("cuda", non_blocking=True),
                     "img_goal_1": host["img_goal_1"].to("cuda", non_blocking=True)},
             "action": host["action"].to("cuda", non_blocking=True)}
        loss = step_p.step(lambda: policy.compute_loss(b))
        loss_host.copy_(loss.reshape(1), non_blocking=True)
    ms_e2e = timed(e2e_step, steps)
    # (c) the same step with the batch already resident
    batch_dev = {"obs": {"img_obs_1": dev["img_obs_1"], "img_goal_1": dev["img_goal_1"]}, "action": dev["action"]}
    ms_p = timed(lambda: step_p.step(lambda: policy.compute_loss(batch_dev)), steps)
    # (d) row N4: the same e2e step fed by the HBM-resident replay buffer (uint8 episodes; per step the host sends a
    # table of 3*B device addresses instead of 2*B float images)
    try:
        import numpy as np
        from v2a_b200.replay import Global_EnvReplayBuffer_Img
        rb = Global_EnvReplayBuffer_Img(["synthetic"], 64, 128, T + 1, None, (128, 128),
                                        env_buf_config={"sample_act_seq_len": T})
        rng = np.random.default_rng(3000 + rank)
        for e in range(32):
            rb.add_one_episode("synthetic", "agentview", e, rng.integers(0, 256, size=(64, 128, 128, 3), dtype=np.uint8),
                               rng.uniform(-1, 1, size=(63, Da)).astype(np.float32))

        def replay_step():
            st, gl, acts, _, _ = rb.sample_random_batch_seq(B)
            b = {"obs": {"img_obs_1": st[:, None], "img_goal_1": gl[:, None]}, "action": acts}   # to_batch_dict
            loss = step_p.step(lambda: policy.compute_loss(b))
            loss_host.copy_(loss.reshape(1), non_blocking=True)
        ms_rb = timed(replay_step, steps)
        replay_e2e = {"value": B * world / (ms_rb * 1e-3), "unit": "samples/s", "ms_per_step": ms_rb,
                      "h2d_bytes_per_step": 3 * B * 8, "d2h_bytes_per_step": 4}
        del rb
    except Exception as exc:   # an optional leg must not cost the bench line
        replay_e2e = {"error": f"{type(exc).__name__}: {exc}"}
    enc_flops, enc_launches = 0.0, 0
    for core in step_p.cores:
        e = OE.last_engine(core)
        enc_flops += sum(gm.flops for gm in e.igemms) + sum(gm.flops for gm in e.wgrads)
        enc_launches += e.planned_launches()
    peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    out = {"metric": POLICY_METRIC, "unit": "samples/s", "batch_per_gpu": B, "steps": steps, "warmup": warm,
           "unet1d_step": {"value": B * world / (ms_u * 1e-3), "ms_per_step": ms_u, "gpu_launches_per_step": int(launches_u),
                           "tensor_flop_per_step": flops, "frac_of_bf16_peak": flops / (ms_u * 1e-3) / 1e12 / peak},
           "compute_loss_step": {"value": B * world / (ms_p * 1e-3), "ms_per_step": ms_p,
                                 "gpu_launches_per_step": int(launches_u) + 2 + int(enc_launches),
                                 "tensor_flop_per_step": flops + enc_flops,
                                 "frac_of_bf16_peak": (flops + enc_flops) / (ms_p * 1e-3) / 1e12 / peak},
           "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": sum(v.numel() * 4 for v in host.values()), "d2h_bytes_per_step": 4},
           "e2e_device_replay": replay_e2e,
           "allreduce": "none (1 rank)" if world == 1 else "NCCL AVG, 64 MiB buckets, each slab started from inside backward"}
    # inference entry: 8-step DDIM predict_action latency at B = 1, device-resident observation
    policy.eval()
    obs1 = {"img_obs_1": dev["img_obs_1"][:1], "img_goal_1": dev["img_goal_1"][:1]}
    with torch.no_grad():
        ms_pa = timed(lambda: policy.predict_action(obs1, use_ddim=True), 10)
    out["predict_action_ms"] = ms_pa
    policy.train()
    step_p.close()
    del step_p, policy
    OE._ENGINES.clear()
    torch.cuda.empty_cache()
    return out


def cpu_baselines(threads):
    """`cpu_baseline` of our line (N = 1): the reference's CPU path on the host cores, bounded samples (~20-30 s)."""
    torch.set_num_threads(threads)
    kind, step = _video_step_fn("cpu")
    g = torch.Generator().manual_seed(0)
    x_cond, te = torch.rand(1, 3, H, W, generator=g), torch.randn(1, TOKENS, 512, generator=g)
    step(x_cond, te)
    t0 = time.perf_counter()
    for _ in range(2):
        step(x_cond, te)
    ts = (time.perf_counter() - t0) / 2
    video = {"value": FRAMES / (ts * DENOISE_STEPS), "unit": UNIT, "cores": threads, "kind": kind,
             "sample": "2 timed denoise steps of ONE video at 128x128x7 through the reference's sample() (1-step DDIM "
                       "plan) after 1 warm-up; a bench step is 100 of these x 16 videos; linear extrapolation"}
    try:
        pk, pstep = _policy_step_fn("cpu", POLICY_B)
        pstep()
        t0 = time.perf_counter()
        pstep()
        ps = time.perf_counter() - t0
        policy = {"value": POLICY_B / ps, "unit": "samples/s", "cores": threads, "kind": pk,
                  "sample": "1 timed compute_loss + backward + clip + AdamW step at B=256 after 1 warm-up"}
    except Exception as exc:
        policy = {"error": f"{type(exc).__name__}: {exc}"}
    return video, policy


def run_ours(args):
    import torch.distributed as dist
    from v2a_b200 import ops
    from v2a_b200.goal_diffusion import GoalGaussianDiffusion
    from v2a_b200.unet import Unet_Libero

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the v2a_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    peaks, peak_src = read_peaks()
    t_start = time.time()

    torch.manual_seed(0)
    net = perturb_(Unet_Libero(), 2)
    diff = GoalGaussianDiffusion(net, image_size=(H, W), channels=3 * FRAMES, timesteps=DENOISE_STEPS,
                                 sampling_timesteps=DENOISE_STEPS, loss_type="l2", objective="pred_v",
                                 beta_schedule="cosine", min_snr_loss_weight=True, guidance_weight=0).cuda()
    g = torch.Generator().manual_seed(1000 + rank)
    cond_host = torch.rand(B, 3, H, W, generator=g).pin_memory()
    te_host = torch.randn(B, TOKENS, 512, generator=g).pin_memory()
    out_host = torch.empty(B, 3 * FRAMES, H, W).pin_memory()
    cond_dev, te_dev = cond_host.cuda(), te_host.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    torch.manual_seed(1234 + rank)
    for _ in range(args.warmup):
        diff.sample(cond_dev, te_dev, batch_size=B)

    # ---- timed: device-resident inputs, EXACTLY K steps ----
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        res = diff.sample(cond_dev, te_dev, batch_size=B)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    frames = args.steps * B * FRAMES * world
    value = frames / (ms * 1e-3)

    # ---- timed: end to end through the public API with host buffers (H2D of the prompts, D2H of the videos) ----
    e2e_steps = max(2, min(args.steps, 4))
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    for _ in range(e2e_steps):
        c = cond_host.to("cuda", non_blocking=True)
        te = te_host.to("cuda", non_blocking=True)
        r = diff.sample(c, te, batch_size=B)
        out_host.copy_(r, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3)) / e2e_steps
    e2e_value = B * FRAMES * world / (ms_e2e * 1e-3)

    eng = net.unet.engine(B, FRAMES, H, W, "cuda")
    kernels_per_denoise = eng_launches_per_step(eng)
    roof = None
    if rank == 0:
        roof = measure_kernel_roofline(diff, B, peaks)
        roof["peak_source"] = peak_src
    # ---- the opt-in one-pass numerics class, same call, N = 1 only (never the headline) ----
    fast = None
    if world == 1 and not args.no_extras:
        try:
            diff.precision = "fast"
            diff.sample(cond_dev, te_dev, batch_size=B)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record()
            for _ in range(2):
                r = diff.sample(cond_host.to("cuda", non_blocking=True), te_host.to("cuda", non_blocking=True), batch_size=B)
                out_host.copy_(r, non_blocking=True)
            f1.record()
            torch.cuda.synchronize()
            fast = {"precision": "fast (one bf16 product per contraction, ~1e-2: the class of the reference's fp16 "
                                 "autocast path; opt-in attribute, own tolerance 2e-2 in tests)",
                    "e2e_value": 2 * B * FRAMES / (f0.elapsed_time(f1) * 1e-3), "unit": UNIT,
                    "ms_per_step": f0.elapsed_time(f1) / 2}
        except Exception as exc:
            fast = {"error": f"{type(exc).__name__}: {exc}"}
        finally:
            diff.precision = "strict"
    out_checksum = float(res.double().mean().item())
    # free the video engines before the policy / baseline legs
    del diff, res, r
    from v2a_b200 import unet as U
    U._ENGINES.clear()
    del eng
    torch.cuda.empty_cache()

    policy_line = None
    if not args.no_policy:
        policy_line = run_policy(args, world, rank, barrier, max_over_ranks, peaks)
    online = None
    if not args.no_extras:
        try:     # BASELINE.json configs[4]: the online loop (video sample + rollout + replay + policy updates) on N GPUs
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import online_loop
            online = online_loop.run_loop(SimpleNamespace(tasks=8, iters=1, policy_steps=10, batch=POLICY_B,
                                                          denoise_steps=DENOISE_STEPS, exec_steps=4, batch_videos=True))
            if online is not None:
                online = {k: online[k] for k in ("n_gpus", "tasks", "batch_videos", "explore_ms_per_iter", "video_frames_per_s",
                                                 "predict_action_calls_per_task", "train_ms_per_step",
                                                 "policy_samples_per_s", "loss", "stubs")}
            from v2a_b200 import obs_encoder as OE
            OE._ENGINES.clear()
            U._ENGINES.clear()
            torch.cuda.empty_cache()
        except Exception as exc:
            online = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        cpu, cpu_policy, eager = None, None, None
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported at N = 1 only
            cpu, cpu_policy = cpu_baselines(os.cpu_count() or 1)
        if world == 1 and not args.no_extras:         # the reference's PyTorch-eager GPU path, same box, same run
            budget = max(60.0, min(240.0, 800.0 - (time.time() - t_start)))
            eager = gpu_eager_subprocess(B, budget)
            try:
                ev, ep = eager["video"], eager["policy"]
                eager["ours_over_eager"] = {
                    "video_e2e_vs_E32": e2e_value / ev["E32"]["frames_per_s"],
                    "video_e2e_vs_E16": e2e_value / ev["E16"]["frames_per_s"],
                    "video_fast_e2e_vs_E16": (fast["e2e_value"] / ev["E16"]["frames_per_s"]) if fast and "e2e_value" in fast else None,
                    "policy_step_vs_E32": policy_line["compute_loss_step"]["value"] / ep["E32"]["samples_per_s"] if policy_line else None,
                    "policy_step_vs_E16": policy_line["compute_loss_step"]["value"] / ep["E16"]["samples_per_s"] if policy_line else None}
            except (KeyError, TypeError):
                pass
        if policy_line is not None and cpu_policy is not None:
            policy_line["cpu_baseline"] = cpu_policy
        summary = {"video_fps": round(value, 3), "video_e2e_fps": round(e2e_value, 3),
                   "ms_per_denoise_step": round(ms / args.steps / DENOISE_STEPS, 3),
                   "igemm_frac_of_bf16_peak": round(roof["frac"], 4),
                   "whole_step_frac_of_bf16_peak": round(B * world * DENOISE_STEPS * FLOP_PER_VIDEO_STEP * args.steps / (ms * 1e-3) / 1e12 /
                                                         (roof["peak"] * world), 4)}
        if policy_line is not None:
            summary["policy_scaling"] = {
                "unet1d_ms": round(policy_line["unet1d_step"]["ms_per_step"], 3), "unet1d_sps": round(policy_line["unet1d_step"]["value"]),
                "loss_ms": round(policy_line["compute_loss_step"]["ms_per_step"], 3), "loss_sps": round(policy_line["compute_loss_step"]["value"]),
                "e2e_ms": round(policy_line["e2e"]["ms_per_step"], 3), "e2e_sps": round(policy_line["e2e"]["value"]),
                "replay_e2e_ms": round(policy_line["e2e_device_replay"].get("ms_per_step", 0.0), 3),
                "predict_action_ms": round(policy_line["predict_action_ms"], 3)}
        if eager is not None and "ours_over_eager" in eager:
            summary["x_eager"] = {k: (round(v, 2) if v else v) for k, v in eager["ours_over_eager"].items()}
        if online is not None and "error" not in online:
            summary["online_loop"] = {"n_gpus": online["n_gpus"], "video_fps": round(online["video_frames_per_s"], 2),
                                      "train_ms": round(online["train_ms_per_step"], 2), "policy_sps": round(online["policy_samples_per_s"])}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3->f32", "data": "synthetic (seeded random-init Unet_Libero weights, "
                "random prompts)", "config": workload_config(world, B),
                "policy_scaling": summary.get("policy_scaling"),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": cond_host.numel() * 4 + te_host.numel() * 4,
                        "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e, "timed_steps": e2e_steps},
                "gpu_launches": int(args.steps * DENOISE_STEPS * roof["launches_per_denoise_step"]),
                "gpu_launches_all_kernels": int(args.steps * (DENOISE_STEPS * kernels_per_denoise + 1)),
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "out_checksum": out_checksum, "policy": policy_line, "gpu_eager": eager, "fast": fast,
                "online_loop": online, "summary": summary}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def eng_launches_per_step(eng):
    """Kernel launches of ours per denoise step: every planned step, 4 launches of the embedding path (counted as one
    step entry), the stats clear and the sampler update."""
    return len(eng.steps) + 3 + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="videos per GPU (configs[1]: 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-policy", action="store_true", help="skip the policy-samples/s object")
    ap.add_argument("--no-extras", action="store_true", help="skip the gpu_eager / fast / online_loop legs")
    ap.add_argument("--policy-steps", type=int, default=20)
    ap.add_argument("--reference-device", default="cpu", choices=["cpu", "cuda"],
                    help="with --impl reference: cuda = the GPU-eager baselines E32/E16 (SURVEY.md §8d)")
    args = ap.parse_args()
    if args.impl == "reference" and args.reference_device == "cuda":
        run_reference_gpu_eager(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
