"""Achieved HBM bandwidth of the bandwidth-bound kernels of the path, at the benchmarked sizes:
fused CFG + DDPM step (forward chunk / reverse step), x_t sampler, GroupNorm apply (streaming, B=100 forward chunk),
LayerNorm, STFT-mel, the late (narrow) vocoder stage's activation pass, layout movers.

Each kernel is timed alone with CUDA events on the launching stream, L2 flushed between iterations (256 MiB write),
median of `--iters`; achieved = ALGORITHMIC bytes (inputs read once + outputs written once) / time, against the
measured copy bandwidth of MEASURED_PEAKS.json (fallback 6552 GB/s).  Under `ncu --profile-from-start off` the same
launches (one each, inside the profiler window) give dram__bytes for the traffic column; see profiles/.

    python tools/hbm_kernels.py [--iters 20] [--json out.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audioeditingcode_b200 import models, unet_config as UC       # noqa: E402
from audioeditingcode_b200.ops import CudaOps                     # noqa: E402

F32, BF16 = torch.float32, torch.bfloat16


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6552.0, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = CudaOps()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, peak_src = peak_gbs()
    rows = []

    def bench(name, fn, bytes_alg, note=""):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        gbs = bytes_alg / us * 1e-3
        rows.append(dict(kernel=name, us=round(us, 2), algorithmic_MB=round(bytes_alg / 1e6, 3), achieved_GBs=round(gbs, 1),
                         frac_of_peak=round(gbs / peak, 3), note=note))
        print(f"{name:44s} {us:9.2f} us  {bytes_alg / 1e6:9.2f} MB  {gbs:8.1f} GB/s  {gbs / peak:6.3f} of peak  {note}")
        # one launch inside the profiler window (a no-op unless running under ncu --profile-from-start off)
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ---------------- scheduler kernels (BASELINE configs[1]: N=200, latent [1,8,256,16])
    N, n_el = 200, 8 * 256 * 16
    cfg = UC.preset("tiny-audioldm")
    from oracle import unet_torch as U
    m = models.load_model("synthetic/audioldm-tiny", dev, N, weights=U.synthetic_weights(cfg, 0), config=cfg)
    x0 = torch.randn(1, 8, 256, 16, device=dev)
    noise = torch.randn(N, 8, 256, 16, device=dev)
    bench("sample_xts_kernel N=200", lambda: m.sample_xts_from_x0(x0, N, noise=noise), (1 + N + N + 1) * n_el * 4,
          "reads x0 + noise[N], writes xts[N+1]")
    xts = m.sample_xts_from_x0(x0, N, noise=noise)
    zs = torch.zeros(N, 8, 256, 16, device=dev)
    cnt = 50
    eps = torch.randn(2 * cnt, 8, 256, 16, device=dev)
    cfg_map = torch.full((1, 8, 256, 16), 3.0, device=dev)
    m.sched_table.set_etas([1.0] * N)
    xt_src = xts.clone()
    bench("cfg_inv_step_kernel count=50 P=1", lambda: m.k_cfg_inv_step(100, cnt, 1.0, eps, eps[cnt:], 1, cfg_map, xt_src, xts,
                                                                       zs, True),
          (6 * cnt + 1) * n_el * 4, "reads eps_u, eps_c, xt, xtm1 (+cfg map), writes xtm1, z")
    xt = torch.randn(1, 8, 256, 16, device=dev)
    out = torch.empty_like(xt)
    bench("cfg_rev_step_kernel P=1 (one step)", lambda: m.k_cfg_rev_step(150, 1.0, eps[:1], eps[1:2], 1, cfg_map, xt, zs[10], out),
          6 * n_el * 4, "768 KiB: launch-latency bound, not a bandwidth kernel")

    # ---------------- GroupNorm apply (streaming) and LayerNorm at the B=100 forward-chunk shape of AudioLDM2-large
    B, HW, Cc = 100, 4096, 192
    a = torch.randn(B * HW, 64, device=dev).to(ops.act_dtype)
    wt = (torch.randn(Cc, 64, device=dev) * 0.1).to(ops.act_dtype)
    x = torch.empty(B * HW, Cc, device=dev)
    cs = torch.zeros(B * Cc * 2, dtype=torch.int64, device=dev)
    ops.gemm(a, wt, out_f32=x, colstats=cs, cs_rows=HW)
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    o = torch.empty(B, HW, Cc, device=dev, dtype=ops.act_dtype)
    bench(f"gn_apply_stream (colstats) [{B},{HW},{Cc}]", lambda: ops.groupnorm(x.view(B, HW, Cc), None, gamma, beta, 1e-5, 32,
                                                                              True, o, cs1=cs),
          B * HW * Cc * 6, "reads fp32 once, writes bf16")
    bench(f"gn stats + apply (two-pass) [{B},{HW},{Cc}]", lambda: ops.groupnorm(x.view(B, HW, Cc), None, gamma, beta, 1e-5,
                                                                                32, True, o),
          B * HW * Cc * 10, "reads fp32 twice, writes bf16")
    M, Cl = 100 * 1024, 384
    h = torch.randn(M, Cl, device=dev)
    g2, b2 = torch.ones(Cl, device=dev), torch.zeros(Cl, device=dev)
    o2 = torch.empty(M, Cl, device=dev, dtype=ops.act_dtype)
    bench(f"layernorm_kernel [{M},{Cl}]", lambda: ops.layernorm(h, g2, b2, o2), M * Cl * 6, "reads fp32, writes bf16")
    M2 = 2 * 1024
    h2 = torch.randn(M2, Cl, device=dev)
    o3 = torch.empty(M2, Cl, device=dev, dtype=ops.act_dtype)
    bench(f"layernorm_kernel [{M2},{Cl}] (B=2 reverse step)", lambda: ops.layernorm(h2, g2, b2, o3), M2 * Cl * 6,
          "4.7 MB: latency bound")

    # ---------------- layout movers / activation passes
    xn = torch.randn(100, 8, 256, 16, device=dev)
    xo = torch.empty(100, 256, 16, 8, device=dev)
    bench("nchw_to_nhwc [100,8,256,16]", lambda: ops.nchw_to_nhwc(xn, out_f32=xo), xn.numel() * 8)
    Tv, Cv = 163840, 32
    xv = torch.randn(1, Tv, Cv, device=dev)
    ov = torch.empty(1, Tv, Cv, device=dev, dtype=ops.act_dtype)
    bench(f"leaky_relu_bf16 [{Tv},{Cv}] (last vocoder stage)", lambda: ops.leaky_relu_bf16(xv, 0.1, ov), Tv * Cv * 6,
          "31 MB")
    ot = torch.empty(1, Tv, device=dev)
    xt1 = torch.randn(1, Tv, device=dev)
    bench(f"tanh_f32 [{Tv}]", lambda: ops.tanh(xt1, ot), Tv * 8, "1.3 MB: latency bound")

    # ---------------- STFT + mel (10.24 s @ 16 kHz)
    from audioeditingcode_b200.audio import TacotronSTFT
    st = TacotronSTFT(1024, 160, 1024, 64, 16000, 0, 8000, device=dev)
    wav = (0.5 * torch.rand(1, 163840, device=dev) - 0.25)
    frames = 163840 // 160 + 1
    bench("stft_mel (10.24 s clip)", lambda: st.mel_spectrogram(wav), 163840 * 4 + frames * (513 + 64) * 4,
          "compute-light (dense DFT, 1.1 GMAC); in 640 KiB, out magnitudes + mel")
    if args.json:
        json.dump(dict(peak_GBs=peak, peak_source=peak_src, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
