#!/usr/bin/env python
"""Benchmark of the DDPM-inversion / CFG-denoising hot path (BASELINE.json metric: denoising-steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

One bench "step" = one complete edit job of one clip on each GPU: `n_inv` inversion steps + `tstart` edit steps
(default BASELINE configs[1]: AudioLDM2-large architecture, 10 s clip -> latent [1,8,256,16], 200-step inversion +
tstart=100 edit, cfg 3 / 12, one source and one target prompt, synthetic seeded weights / text embeddings).
1 denoising step = one classifier-free-guided step = 2 U-Net evaluations + CFG combine + scheduler update
(BASELINE.md §3).  value = denoising steps of ALL ranks / max-over-ranks device time.

Keys (see the task contract): value = inputs resident in HBM; e2e = same job through the public wrapper API with
the clip latent in pinned host memory (H2D inside the timed region, edited latent read back D2H); roofline = the
dominant kernel family (tcgen05 GEMM / implicit conv) FLOP rate from a live CUDA-event pass; cpu_baseline = the
oracle port (oracle/unet_torch.py + oracle/ddpm_oracle.py, fp32 torch on the host cores) on a bounded sample.
--impl reference times that oracle port alone (the reference's own diffusers pipeline cannot travel to the GPU box:
no diffusers, no weights, no network — DESIGN.md §oracle).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1]
    "audioldm2-large-10s": dict(preset="audioldm2-large", model_id="cvssp/audioldm2-large", H=256, W=16, n_inv=200,
                                tstart=100, cfg_src=3.0, cfg_tar=12.0, text_lens=(8, 16)),
    # BASELINE.json configs[0] geometry (the reference's CPU-runnable case) — parity-test sized
    "audioldm-s-5s": dict(preset="audioldm-s", model_id="cvssp/audioldm-s-full-v2", H=128, W=16, n_inv=50, tstart=50,
                          cfg_src=1.0, cfg_tar=3.0, text_lens=()),
    # BASELINE.json configs[2] geometry (TANGO-full, v-prediction scheduler of the checkpoint is set by the wrapper), one clip
    "tango-10s": dict(preset="tango", model_id="declare-lab/tango", H=256, W=16, n_inv=200, tstart=100, cfg_src=3.0,
                      cfg_tar=12.0, text_lens=(16,)),
    # BASELINE.json configs[4] geometry: 30 s clip, AudioLDM2 (whole clip on one GPU)
    "audioldm2-30s": dict(preset="audioldm2", model_id="cvssp/audioldm2", H=768, W=16, n_inv=200, tstart=100, cfg_src=3.0,
                          cfg_tar=12.0, text_lens=(8, 16)),
    # BASELINE.json configs[4]: SDEdit (main_run_sdedit.py:78-100), AudioLDM2, 30 s clip (whole clip on one GPU, DESIGN.md
    # §5: overlapping tiles would change GroupNorm / self-attention results), add_noise at timesteps[100] + 100 steps
    "sdedit-30s": dict(preset="audioldm2", model_id="cvssp/audioldm2", H=768, W=16, n_inv=200, tstart=100, cfg_src=3.0,
                       cfg_tar=12.0, text_lens=(8, 16), mode="sdedit"),
    # BASELINE.json configs[3]: unsupervised PC extraction (main_pc_extract_inv.py:199-209) at ONE timestep of the drift
    # window: n_evs = 8 directions, 50 subspace iterations, const 1e-3; under torchrun the directions are sharded over the
    # ranks (one all-gather of the iterate per iteration) -> strong scaling of one extraction
    "pc-drift": dict(preset="audioldm2", model_id="cvssp/audioldm2", H=256, W=16, n_inv=200, tstart=100, cfg_src=3.0,
                     cfg_tar=3.0, text_lens=(8, 16), mode="pc", n_ev=8, iters=50),
    "tiny": dict(preset="tiny-audioldm2", model_id="synthetic/audioldm2-tiny", H=32, W=16, n_inv=20, tstart=10,
                 cfg_src=3.0, cfg_tar=12.0, text_lens=(8, 16)),
}


def workload_config(name, spec, clips):
    """The `config` object of the JSON line — identical for both arms (ours / --impl reference)."""
    return {"workload": name, "arch": spec["preset"], "mode": spec.get("mode", "edit"), "clips_per_gpu": clips,
            "latent": [1, 8, spec["H"], spec["W"]], "n_inv": spec["n_inv"], "tstart": spec["tstart"]}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            j = json.load(fh)
        return dict(hbm_gbs=j.get("hbm_gbs"), tflops=j.get("bf16_tflops_sustained") or j.get("bf16_tflops"),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- ours
def synth_text(cfg, lens, P, device, seed=4):
    g = torch.Generator().manual_seed(seed)
    dims = {s[1]: s[0] for s in cfg.transformer_specs if s is not None}
    streams = [torch.randn(P, lens[i], dims[i], generator=g).to(device) for i in range(cfg.n_streams)]
    return streams


def build_model(spec, device):
    from audioeditingcode_b200 import models, unet_config as C
    cfg = C.preset(spec["preset"])
    m = models.load_model(spec["model_id"], device, spec["n_inv"], config=cfg, allow_synthetic=True)
    return m, cfg


def run_job(m, spec, x0_dev, forward_batch, fwd_group=None):
    """One bench step.  mode "edit" (default): inversion + edit of the clip(s) in x0_dev ([K,C,H,W]; K > 1 -> the
    multi-clip entry points, B = K*(1+P) rows per launch); "sdedit": add_noise + tstart forward_directional steps;
    "pc": one get_eigenvectors call (n_ev directions, `iters` subspace iterations)."""
    from audioeditingcode_b200.ddm_inversion import inversion_utils as IU
    N, ts = spec["n_inv"], spec["tstart"]
    mode = spec.get("mode", "edit")
    src, tgt = "a recording of a dog barking", "a recording of a cat meowing"
    if mode == "edit" and x0_dev.shape[0] == 1:
        _, zs, xts, _ = IU.inversion_forward_process(m, x0_dev, etas=1.0, prompts=[src], cfg_scales=[spec["cfg_src"]],
                                                     num_inference_steps=N, numerical_fix=True, forward_batch=forward_batch,
                                                     group=fwd_group)
        w, _ = IU.inversion_reverse_process(m, xT=xts, tstart=torch.tensor([ts], dtype=torch.int), etas=1.0,
                                            prompts=[tgt], neg_prompts=[""], cfg_scales=[spec["cfg_tar"]], zs=zs[:ts])
        return w
    if mode == "edit":
        _, zs, xts = IU.inversion_forward_process_batched(m, x0_dev, etas=1.0, prompts=[src], cfg_scales=[spec["cfg_src"]],
                                                          num_inference_steps=N, numerical_fix=True,
                                                          forward_batch=forward_batch)
        w, _ = IU.inversion_reverse_process_batched(m, xts, ts, etas=1.0, prompts=[tgt], neg_prompts=[""],
                                                    cfg_scales=[spec["cfg_tar"]], zs=zs[:, :ts])
        return w
    from audioeditingcode_b200 import pc_drift as PC
    st = _pc_state(m, spec, x0_dev, src)
    if mode == "sdedit":                                        # main_run_sdedit.py:89-100
        timesteps = m.model.scheduler.timesteps
        skip = N - ts
        xt = m.model.scheduler.add_noise(x0_dev, st["noise"], timesteps[skip:][:1].unsqueeze(0))
        for it, t in enumerate(timesteps[skip:]):
            xt, _ = PC.forward_directional(m, xt, t, st["latents"][skip + it + 1][None], st["unc"], st["txt"],
                                           spec["cfg_tar"], eta=1)
        return xt
    # mode == "pc": main_pc_extract_inv.py:199-209 at one timestep of the drift window
    t = m.model.scheduler.timesteps[N - ts]
    _, x0p = PC.forward_directional(m, x0_dev, t, st["latents"][0][None], st["unc"], st["txt"], spec["cfg_tar"], eta=1)
    ev = PC.get_eigenvectors(m, x0_dev, st["txt"], st["unc"], st["latents"][0][None], st["mask"], t, x0p,
                             PC.PCStreamChoice.BOTH, 1e-3, spec["cfg_tar"], spec["iters"], False, 1, spec["n_ev"],
                             group=st["group"])
    return ev[0]


_PC_STATE = {}


def _pc_state(m, spec, x0_dev, prompt):
    """Per-model constants of the sdedit / pc jobs (text embeddings, pre-drawn latents), built once outside the timing."""
    st = _PC_STATE.get(id(m))
    if st is None:
        from audioeditingcode_b200 import pc_drift as PC
        g = torch.Generator(device=x0_dev.device).manual_seed(7)
        N = spec["n_inv"]
        unc, txt = m.encode_text([""], negative=True), m.encode_text([prompt])
        group = None
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            group = dist.group.WORLD
        st = dict(unc=PC.PromptEmbeddings(*unc), txt=PC.PromptEmbeddings(*txt), group=group,
                  latents=torch.randn((N + 1, *x0_dev.shape[1:]), device=x0_dev.device, generator=g),
                  noise=torch.randn(x0_dev.shape, device=x0_dev.device, generator=g), mask=torch.ones_like(x0_dev))
        _PC_STATE[id(m)] = st
    return st


def ends_pass(m, spec, reps=3):
    """The ends of the path timed on the device, reported beside the loop number as BASELINE.md §3 asks: log-mel STFT of
    the clip's waveform (ae_stft_mel), VAE encode / decode (tcgen05 convs), HiFi-GAN vocoder (called twice per edit in
    main_run.py:184-185).  Algorithmic FLOPs at 10.24 s from BASELINE.md §2 (scaled linearly with the clip length)."""
    dev = m.device
    T = spec["H"] * 4                                    # mel frames
    scale = T / 1024.0
    wav = (0.5 * torch.rand(1, T * 160, device=dev) - 0.25)
    mel = torch.randn(1, 1, T, 64, device=dev) * 2.0 - 5.0
    lat = torch.randn(1, 8, spec["H"], spec["W"], device=dev) * 0.5
    stft = m.get_fn_STFT()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps
    out = {}
    for name, fn, gflop in (("stft_mel", lambda: stft.mel_spectrogram(wav), 2.2), ("vae_encode", lambda: m.vae_encode(mel), 345.0),
                            ("vae_decode", lambda: m.vae_decode(lat), 636.0),
                            ("vocoder", lambda: m.decode_to_mel(mel), 1030.0)):
        ms = timed(fn)
        out[name] = {"ms": round(ms, 3), "algorithmic_gflop": round(gflop * scale, 1),
                     "tflops": round(gflop * scale / ms, 2)}
    out["per_edit_ms"] = round(out["stft_mel"]["ms"] + out["vae_encode"]["ms"] + out["vae_decode"]["ms"] +
                               2 * out["vocoder"]["ms"], 2)
    out["what"] = "synthetic VAE / vocoder weights; one call each, CUDA events; vocoder runs twice per edit"
    return out


def gemm_event_pass(m, spec, cfg, B=2):
    """Time share and FLOP rate of the dominant kernel family (ae_gemm: tcgen05 GEMM / implicit conv + split-K reduce).
    The ae_gemm calls of one U-Net evaluation (batch B) are recorded, then replayed ALONE, back to back, inside one
    CUDA graph bracketed by CUDA events on the launching stream (no host launch gaps, warm caches); the whole
    evaluation is timed the same way through its own graph."""
    ops = m.engine.ops
    dev = m.device
    from audioeditingcode_b200.ddm_inversion.inversion_utils import _loop_text
    text, cl = _loop_text(m, [""], ["a recording of a dog barking"])
    x = torch.randn(B, cfg.in_channels, spec["H"], spec["W"], device=dev)
    t = torch.full((B,), 501, dtype=torch.int64, device=dev)
    slot = (torch.arange(B, dtype=torch.int32, device=dev) % 2) if text is not None else None
    clb = None if cl is None else cl[(torch.arange(B, device=dev) % 2)]
    calls, keep = [], []
    orig = ops.gemm
    orig_empty = ops.empty

    def rec(*a, **k):
        calls.append((a, k))
        orig(*a, **k)

    def keep_empty(*a, **k):
        tns = orig_empty(*a, **k)
        keep.append(tns)          # keep every operand alive so the recorded pointers stay valid
        return tns
    m.engine.forward(x, t, text=text, slot_map=slot, class_labels=clb)
    ops.gemm, ops.empty = rec, keep_empty
    m.engine.forward(x, t, text=text, slot_map=slot, class_labels=clb)
    ops.gemm, ops.empty = orig, orig_empty
    torch.cuda.synchronize()

    def time_graph(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    l0 = ops.launch_count()
    gemm_ms = time_graph(lambda: [orig(*a, **k) for a, k in calls])
    n_launch = (ops.launch_count() - l0) // 2
    eval_ms = time_graph(lambda: m.engine.forward(x, t, text=text, slot_map=slot, class_labels=clb))
    return dict(gemm_ms=gemm_ms, eval_ms=eval_ms, n_gemm=len(calls), n_gemm_launches=n_launch)


def flops_per_eval(cfg, spec, B):
    from audioeditingcode_b200.flops import count_flops
    return count_flops(cfg, spec["H"], spec["W"], B, spec["text_lens"])


def usable_cores():
    """Host cores this process may actually run on: scheduler affinity capped by the cgroup CPU quota (a container
    that sees 128 CPUs but is throttled to a few would otherwise be oversubscribed by torch's thread pool)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def pick_threads(limit, probe=None):
    """Thread count for the CPU arm, calibrated ON THE WORKLOAD ITSELF: one U-Net evaluation of the oracle port (`probe`)
    per candidate count <= the usable cores, after a warm-up evaluation; the fastest wins (more threads than the machine
    really grants is slower, not faster).  Round 1 calibrated on a lone conv3x3 and picked 8 or 16 threads on the same
    box class (0.90 vs 1.28 steps/s); the evaluation-level probe is what the arm then runs."""
    cands = sorted({c for c in (limit, limit // 2, 16, 8) if 1 <= c <= limit}, reverse=True)
    if probe is None or len(cands) == 1:
        torch.set_num_threads(cands[0])
        return cands[0]
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        probe()
        t0 = time.perf_counter()
        probe()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


class CpuOracle:
    """The oracle port (oracle/unet_torch.py + oracle/ddpm_oracle.py, fp32 torch CPU) set up once for the bench
    workload; `steps(n)` runs n CFG denoising steps (alternating inversion / edit updates) and returns seconds."""

    def __init__(self, spec, threads):
        from oracle import unet_torch as U
        from oracle import ddpm_oracle as D
        from audioeditingcode_b200 import unet_config as C
        torch.set_num_threads(threads)
        self.U, self.D, self.spec = U, D, spec
        cfg = self.cfg = C.preset(spec["preset"])
        self.w = U.synthetic_weights(cfg, seed=0)
        g = torch.Generator().manual_seed(1)
        H, Wd = spec["H"], spec["W"]
        x0 = 0.5 * torch.randn(1, cfg.in_channels, H, Wd, generator=g)
        dims = {s[1]: s[0] for s in cfg.transformer_specs if s is not None}
        lens = spec["text_lens"]
        self.streams_u = [torch.randn(1, 1 if i == cfg.n_streams - 1 else lens[i], dims[i], generator=g)
                          for i in range(cfg.n_streams)]
        self.streams_c = [torch.randn(1, lens[i], dims[i], generator=g) for i in range(cfg.n_streams)]
        self.yu = torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1) if cfg.class_embed_dim else None
        self.yc = torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1) if cfg.class_embed_dim else None
        self.sched = D.MiniDDIM(cfg.beta_start, cfg.beta_end, prediction_type=cfg.prediction_type)
        self.sched.set_timesteps(spec["n_inv"])
        self.N = spec["n_inv"]
        self.noise = torch.randn(self.N, *x0.shape[1:], generator=g)
        self.xts = D.sample_xts_from_x0(self.sched, x0, self.noise)
        self.cfgm, _ = D.build_cfg_maps(1, x0.shape[1:], [spec["cfg_src"]], None)
        self.pos = 0

    def unet(self, x, t, which):
        tt = torch.full((x.shape[0],), int(t), dtype=torch.int64)
        st = self.streams_u if which == "uncond" else self.streams_c
        with torch.no_grad():
            return self.U.unet_forward(self.cfg, self.w, x, tt, streams=st, stream_masks=[None] * len(st),
                                       class_labels=(self.yu if which == "uncond" else self.yc))[0]

    def one_step(self):
        D, pos, N = self.D, self.pos % (self.N - 1), self.N
        self.pos += 1
        t = int(self.sched.timesteps[pos])
        idx = N - pos - 1
        xt = self.xts[idx + 1][None]
        eps = D.cfg_combine(self.unet(xt, t, "uncond"), self.unet(xt, t, "cond"), self.cfgm)
        if pos % 2 == 0:
            D.get_zs_from_xts(self.sched, xt, self.xts[idx][None], eps, t, 1.0, True)
        else:
            D.reverse_step_with_custom_noise(self.sched, eps, t, xt, self.noise[idx][None], 1.0)

    def steps(self, n):
        t0 = time.perf_counter()
        for _ in range(n):
            self.one_step()
        return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="audioldm2-large-10s", choices=sorted(CONFIGS))
    ap.add_argument("--forward-batch", type=int, default=int(os.environ.get("AEDIT_FORWARD_BATCH", "50")))
    ap.add_argument("--clips", type=int, default=1, help="clips per GPU per job (K > 1: multi-clip entry points, "
                    "B = K*(1+P) rows per U-Net launch; BASELINE configs[2] uses 4)")
    ap.add_argument("--queue-group", type=int, default=4, help="clips per reverse launch of the extra throughput_queue "
                    "measurement (0/1 = skip it)")
    ap.add_argument("--cpu-steps", type=int, default=0, help="CFG steps of the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ends", action="store_true", help="skip the STFT / VAE / vocoder timing")
    ap.add_argument("--shard-forward", action="store_true",
                    help="N > 1: strong scaling of ONE clip — the timestep chunks of its forward process are sharded over "
                         "the ranks (all-gather of the owned zs / xts rows), every rank then runs the sequential reverse "
                         "process (replicas); default is one clip job per rank (weak)")
    args = ap.parse_args()
    spec = CONFIGS[args.config]
    mode = spec.get("mode", "edit")
    K = max(1, args.clips) if mode == "edit" else 1
    # denoising steps (CFG-guided U-Net steps) one bench step performs on one GPU
    if mode == "edit":
        steps_per_job = (spec["n_inv"] + spec["tstart"]) * K
    elif mode == "sdedit":
        steps_per_job = spec["tstart"]
    else:                                    # pc: 1 unperturbed step + iters iterations of n_ev perturbed CFG steps
        steps_per_job = 1 + spec["iters"] * spec["n_ev"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = usable_cores()

    # ------------------------------------------------------------------------------ reference arm (CPU oracle port)
    if args.impl == "reference":
        if rank != 0:
            return
        oracle = CpuOracle(spec, cores)
        threads = pick_threads(cores, lambda: oracle.unet(oracle.xts[1][None], 1, "uncond"))
        oracle.unet(oracle.xts[1][None], 1, "uncond")     # one untimed evaluation: thread pool / primitive caches
        n_sample = args.cpu_steps or 1
        vals = []
        for i in range(args.warmup + args.steps):
            dt = oracle.steps(n_sample)
            if i >= args.warmup:
                vals.append((n_sample / dt, dt))
        cores = threads
        tot_steps = n_sample * len(vals)
        tot_t = sum(dt for _, dt in vals)
        value = tot_steps / tot_t
        line = {"metric": "denoising-steps/sec", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot_t / max(1, len(vals)),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "impl": "reference",
                "config": workload_config(args.config, spec, K),
                "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                                 "sample": f"{n_sample} CFG denoising steps (2 U-Net evals each, fp32 torch CPU oracle port) "
                                           f"per bench step of the {args.config} workload"},
                "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    m, cfg = build_model(spec, dev)
    ops = m.engine.ops
    shard_fwd = args.shard_forward and world > 1 and mode == "edit" and K == 1
    fwd_group = dist.group.WORLD if shard_fwd else None
    g = torch.Generator().manual_seed(1 if shard_fwd else 1 + rank)       # sharded: every rank holds the same clip
    x0_host = (0.5 * torch.randn(K, cfg.in_channels, spec["H"], spec["W"], generator=g)).pin_memory()
    x0_dev = x0_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > L2 (126 MB)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, K):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(K):
            flush.zero_()
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def job_resident():
        run_job(m, spec, x0_dev, args.forward_batch, fwd_group)

    out_rows = spec["n_ev"] if mode == "pc" else K
    e2e_out = torch.empty(out_rows, cfg.in_channels, spec["H"], spec["W"]).pin_memory()

    def job_e2e():
        x = x0_host.to(dev, non_blocking=True)
        w = run_job(m, spec, x, args.forward_batch, fwd_group)
        e2e_out.copy_(w, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(3, args.warmup)):
        job_resident()
    with ClockSampler(local_rank) as cs:
        l0 = ops.launch_count() + getattr(m, "graph_kernels", 0)
        ms = timed_loop(job_resident, args.steps)
        launches = ops.launch_count() + getattr(m, "graph_kernels", 0) - l0   # eager launches + kernels replayed in graphs
    clocks = cs.summary()
    ms_e2e = timed_loop(job_e2e, args.steps)
    # pc under torchrun shards ONE extraction over the ranks (strong scaling); everything else is one job per rank (weak)
    strong = (mode == "pc" and world > 1) or shard_fwd
    total_steps = steps_per_job * args.steps * (1 if strong else world)
    from audioeditingcode_b200 import _lib as _aelib
    op_dtype = "fp16" if _aelib.load().ae_operand_dtype() == 1 else "bf16"
    value = total_steps / (ms / 1000.0)
    e2e_value = total_steps / (ms_e2e / 1000.0)

    line = {"metric": "denoising-steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": op_dtype, "data": "synthetic",
            "config": workload_config(args.config, spec, K),
            "details": {"weights": m.weights_source,
                        "precision": f"{op_dtype} tensor-core operands (AEDIT_OPERANDS), fp32 accumulate / residual stream / "
                                     "scheduler state",
                        "denoising_steps_per_bench_step": steps_per_job, "forward_batch_timesteps": args.forward_batch,
                        "parallelism": (f"forward-timesteps-sharded-{world}+reverse-replicated" if shard_fwd else
                                        f"pc-directions-sharded-{world}" if strong else f"clip-dp{world}"),
                        "l2": "256 MiB flush between jobs; weights (1.5 GB) >> L2",
                        "lanes": ("forward chunks and reverse steps of the clip on two streams (reverse lane high priority), "
                                  f"fast path taken {getattr(m, 'overlap_hits', 0)}x") if getattr(m, "overlap_hits", 0)
                        else "single stream"},
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": x0_host.numel() * 4,
                    "d2h_bytes_per_step": e2e_out.numel() * 4},
            "gpu_launches": int(launches), "clocks": clocks}

    # Serving-style throughput of the SAME workload with several clips in flight per GPU (N = 1 only): a queue of
    # 4*Q clips edited with Q clips per reverse launch (B = 2Q rows) while the next group's forward process runs on the
    # forward lane (inversion_utils.edit_clips_pipelined).  Reported beside `value` (one clip at a time), never instead.
    if rank == 0 and world == 1 and mode == "edit" and K == 1 and args.queue_group > 1:
        from audioeditingcode_b200.ddm_inversion import inversion_utils as IU
        Q = args.queue_group
        NG = 4                       # groups in the queue: 1 fill + 3 steady-state periods (forward(g+1) || reverse(g)) + drain
        gq = torch.Generator().manual_seed(99)
        q_host = [(0.5 * torch.randn(1, cfg.in_channels, spec["H"], spec["W"], generator=gq)).pin_memory()
                  for _ in range(NG * Q)]
        q_out = [torch.empty(1, cfg.in_channels, spec["H"], spec["W"]).pin_memory() for _ in range(NG * Q)]

        def queue_job():
            IU.edit_clips_pipelined(m, q_host, ["a recording of a dog barking"], ["a recording of a cat meowing"],
                                    spec["tstart"], cfg_src=spec["cfg_src"], cfg_tar=spec["cfg_tar"],
                                    num_inference_steps=spec["n_inv"], forward_batch=args.forward_batch, group=Q,
                                    on_result=lambda i, w: q_out[i].copy_(w, non_blocking=True))
            torch.cuda.synchronize()
        queue_job()                                              # graph captures for the B = 2Q shapes
        ms_q = timed_loop(queue_job, 1)
        line["throughput_queue"] = {
            "value": (spec["n_inv"] + spec["tstart"]) * NG * Q / (ms_q / 1000.0), "unit": "steps/s",
            "clips_in_flight": Q, "clips": NG * Q, "ms_per_clip": ms_q / (NG * Q),
            "what": "same workload, host-resident clips (H2D / D2H inside the timed region): groups of Q clips per "
                    "reverse launch, the next group's forward process overlapped on the forward lane"}
    if rank == 0 and (mode != "edit" or K > 1):
        # other workloads: one evaluation shape dominates; report its GEMM-family rate only
        peaks = load_peaks()
        rows = 2 * (spec["n_ev"] if mode == "pc" else K)
        fl = flops_per_eval(cfg, spec, rows)
        gp = gemm_event_pass(m, spec, cfg, rows)
        ach = (fl["conv"] + fl["linear"]) / gp["gemm_ms"] / 1e9
        line["roofline"] = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (conv3x3 / conv1x1 / linear)",
                            "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                            "traffic": None, "peak_source": peaks["source"],
                            "per_eval": {f"B{rows}": {"gemm_ms": gp["gemm_ms"], "eval_ms": gp["eval_ms"]}}}
        print(json.dumps(line))
    elif rank == 0:
        peaks = load_peaks()
        fl2 = flops_per_eval(cfg, spec, 2)
        gp = gemm_event_pass(m, spec, cfg, 2)
        # dominant kernel over the job: the forward-process chunks of the plan the loop actually runs (B = 2 x
        # timesteps per chunk) + tstart reverse steps at B = 2
        from collections import Counter
        from audioeditingcode_b200.ddm_inversion.inversion_utils import _chunk_plan
        plan = Counter(c for _, c in _chunk_plan(spec["n_inv"], args.forward_batch, spec["n_inv"] // 2))
        gemm_flops = spec["tstart"] * (fl2["conv"] + fl2["linear"])
        gemm_ms_job = spec["tstart"] * gp["gemm_ms"]
        eval_ms_job = spec["tstart"] * gp["eval_ms"]
        per_chunk = {}
        for count, n_c in sorted(plan.items()):
            flc = flops_per_eval(cfg, spec, 2 * count)
            gpc = gemm_event_pass(m, spec, cfg, 2 * count)
            gemm_flops += n_c * (flc["conv"] + flc["linear"])
            gemm_ms_job += n_c * gpc["gemm_ms"]
            eval_ms_job += n_c * gpc["eval_ms"]
            per_chunk[f"B{2 * count}"] = {"chunks": n_c, "gemm_ms": gpc["gemm_ms"], "eval_ms": gpc["eval_ms"],
                                          "tflops": (flc["conv"] + flc["linear"]) / gpc["gemm_ms"] / 1e9}
        achieved = gemm_flops / (gemm_ms_job / 1000.0) / 1e12
        # whole-job FLOPs: forward batches B = 2*forward_batch per launch, reverse B = 2
        job_flops = (spec["n_inv"] + spec["tstart"]) * fl2["total"]
        traffic, traffic_src, alg_bytes = None, None, None
        for tname in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
            tj = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", tname)
            if os.path.exists(tj):  # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture
                tinfo = json.load(open(tj))
                traffic, traffic_src = tinfo["traffic_bytes_per_launch"], tinfo["source"]
                alg_bytes = tinfo.get("algorithmic_bytes_per_launch")
                break
        in_situ = gemm_flops / (ms / args.steps / 1000.0) / 1e12      # same FLOPs over the whole job time (all kernels)
        line["roofline"] = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (conv3x3 / conv1x1 / linear)",
                            "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                            "frac": achieved / peaks["tflops"], "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": alg_bytes,
                            "achieved_in_situ": in_situ, "frac_in_situ": in_situ / peaks["tflops"],
                            "how": "achieved = conv+linear FLOPs of the job (SURVEY 8d counting rule, the reference "
                                   "algorithm's layers) / time of the job's ae_gemm launches replayed ALONE back to back in a "
                                   "CUDA graph; achieved_in_situ = the same FLOPs / ms_per_step (all kernels, both lanes)",
                            "peak_source": peaks["source"],
                            "flops_per_job": gemm_flops, "gemm_ms_per_job": gemm_ms_job,
                            "gemm_share_of_unet_time": gemm_ms_job / eval_ms_job,
                            "unet_time_serial_over_job_time": eval_ms_job / (ms / args.steps),
                            "per_eval": dict({"B2": {"gemm_ms": gp["gemm_ms"], "eval_ms": gp["eval_ms"],
                                                     "gemm_calls": gp["n_gemm"],
                                                     "tflops": (fl2["conv"] + fl2["linear"]) / gp["gemm_ms"] / 1e9}},
                                             **per_chunk),
                            "job_tflops_all_kernels": job_flops * args.steps / (ms / 1000.0) / 1e12}
        if world == 1 and not args.no_ends:
            try:
                line["ends"] = ends_pass(m, spec)
            except Exception as ex:             # the loop numbers must not be lost to an ends problem
                line["ends"] = {"error": repr(ex)[:200]}
        if not args.no_cpu_baseline and world == 1:
            oracle = CpuOracle(spec, cores)
            threads = pick_threads(cores, lambda: oracle.unet(oracle.xts[1][None], 1, "uncond"))
            oracle.unet(oracle.xts[1][None], 1, "uncond")     # untimed warm-up evaluation
            n_sample = args.cpu_steps or 1
            dt = oracle.steps(n_sample)
            if dt < 10.0 and not args.cpu_steps:               # fast host: extend the sample to ~15 s
                extra = min(20, int(15.0 / (dt / n_sample)))
                dt += oracle.steps(extra)
                n_sample += extra
            line["cpu_baseline"] = {"value": n_sample / dt, "unit": "steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_sample} CFG denoising step(s) of the same workload on the fp32 torch "
                                              f"CPU oracle port, {threads} threads of {cores} usable cores ({dt:.1f} s)"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
