"""Probe: how fast does the reverse-lane graph (B=2, latency-bound chain) run while a forward-chunk graph (B=100,
throughput-bound) occupies the machine — plain priority streams vs SM partitions (green contexts)?"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from audioeditingcode_b200.ddm_inversion import inversion_utils as IU  # noqa: E402


def main():
    spec = B.CONFIGS["audioldm2-large-10s"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m, cfg = B.build_model(spec, dev)
    x0 = (0.5 * torch.randn(1, cfg.in_channels, spec["H"], spec["W"], generator=torch.Generator().manual_seed(1))).to(dev)
    IU.OVERLAP = True
    IU.REV_VARIANT = "solo"
    B.run_job(m, spec, x0, 50)
    torch.cuda.synchronize()
    fwd = rev = None
    for k, g in m.engine._graphs.items():
        if k[4] == ("fwd", 50, 1) and k[-1] == 0:
            fwd = g
        if k[4] == ("rev", 1) and k[-1] == 1:
            rev = g
    assert fwd is not None and rev is not None, list(m.engine._graphs)

    def t_alone(g, stream, reps):
        with torch.cuda.stream(stream):
            g.graph.replay()
            stream.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            for _ in range(reps):
                g.graph.replay()
            e.record(stream)
            stream.synchronize()
        return s.elapsed_time(e) / reps

    def concurrent(sf, sr, n_fwd=2):
        nonlocal fwd, rev
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sf):
            s.record(sf)
            for _ in range(n_fwd):
                fwd.graph.replay()
            e.record(sf)
        n = 0
        evs = []
        t0 = time.perf_counter()
        with torch.cuda.stream(sr):
            while not e.query():
                if len(evs) >= 2:
                    evs.pop(0).synchronize()
                rev.graph.replay()
                ev = torch.cuda.Event()
                ev.record(sr)
                evs.append(ev)
                n += 1
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n_fwd, n / n_fwd

    s0 = torch.cuda.Stream()
    shi = torch.cuda.Stream(priority=-1)
    r_alone = t_alone(rev, shi, 10)
    f_alone = t_alone(fwd, s0, 2)
    print(json.dumps({"rev_alone_ms": round(r_alone, 3), "fwd_alone_ms": round(f_alone, 2)}), flush=True)
    f_ms, n = concurrent(s0, shi)
    print(json.dumps({"mode": "priority streams", "fwd_ms": round(f_ms, 2), "rev_steps_per_fwd": n,
                      "rev_ms_per_step": round(f_ms / max(n, 1), 2)}), flush=True)
    f_ms, n = concurrent(s0, torch.cuda.Stream())
    print(json.dumps({"mode": "plain streams", "fwd_ms": round(f_ms, 2), "rev_steps_per_fwd": n,
                      "rev_ms_per_step": round(f_ms / max(n, 1), 2)}), flush=True)

    # ---- green contexts
    from cuda.bindings import driver as drv

    def chk(r):
        if r[0] != drv.CUresult.CUDA_SUCCESS:
            raise RuntimeError(str(r[0]))
        return r[1:] if len(r) > 2 else r[1]

    cudev = chk(drv.cuDeviceGet(0))
    sm = chk(drv.cuDeviceGetDevResource(cudev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    print(json.dumps({"device_sms": sm.sm.smCount}), flush=True)
    from audioeditingcode_b200.unet import GraphedForward
    Bq, _, Hq, Wq = rev.x.shape
    for n_small in (32, 48, 64):
        try:
            groups, nb, rem = chk(drv.cuDevSmResourceSplitByCount(1, sm, 0, n_small))
            d_small = chk(drv.cuDevResourceGenerateDesc([groups[0]], 1))
            d_rem = chk(drv.cuDevResourceGenerateDesc([rem], 1))
            g_small = chk(drv.cuGreenCtxCreate(d_small, cudev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
            g_rem = chk(drv.cuGreenCtxCreate(d_rem, cudev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
            st_small = chk(drv.cuGreenCtxStreamCreate(g_small, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, -5))
            st_rem = chk(drv.cuGreenCtxStreamCreate(g_rem, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
            ts, tr = torch.cuda.ExternalStream(int(st_small)), torch.cuda.ExternalStream(int(st_rem))
            info = {"green_small_sms": groups[0].sm.smCount, "green_rem_sms": rem.sm.smCount}
            # capture the graphs INSIDE the green contexts (a graph instantiated in the primary context ignores the
            # partition of the stream it is launched into: measured)
            ctx_s = chk(drv.cuCtxFromGreenCtx(g_small))
            (err,) = drv.cuCtxPushCurrent(ctx_s)
            assert err == drv.CUresult.CUDA_SUCCESS, err
            try:
                rev_g = GraphedForward(m.engine, Bq, Hq, Wq, rev.text, rev.slot, rev.cl, lane=1, capture_stream=ts)
            finally:
                drv.cuCtxPopCurrent()
            info["rev_alone_small_ms"] = round(t_alone(rev_g, ts, 10), 3)
            same = bool(torch.equal(rev_g(rev.x, rev.t, rev.cl), rev(rev.x, rev.t, rev.cl)))
            info["same_bits"] = same
            old_rev = rev
            rev = rev_g
            try:
                f_ms, n = concurrent(s0, ts)
                info.update({"fwd_all_sms_ms": round(f_ms, 2), "rev_steps_per_fwd_all": n,
                             "rev_ms_per_step_all": round(f_ms / max(n, 1), 2)})
                ctx_r = chk(drv.cuCtxFromGreenCtx(g_rem))
                (err,) = drv.cuCtxPushCurrent(ctx_r)
                assert err == drv.CUresult.CUDA_SUCCESS, err
                try:
                    fwd_g = GraphedForward(m.engine, fwd.x.shape[0], Hq, Wq, fwd.text, fwd.slot, fwd.cl, lane=0,
                                           capture_stream=tr)
                finally:
                    drv.cuCtxPopCurrent()
                info["fwd_alone_rem_ms"] = round(t_alone(fwd_g, tr, 2), 2)
                old_fwd = fwd
                fwd = fwd_g
                try:
                    f_ms, n = concurrent(tr, ts)
                    info.update({"fwd_ms": round(f_ms, 2), "rev_steps_per_fwd": n,
                                 "rev_ms_per_step": round(f_ms / max(n, 1), 2)})
                finally:
                    fwd = old_fwd
            finally:
                rev = old_rev
            print(json.dumps(info), flush=True)
        except Exception as ex:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(json.dumps({"green_small": n_small, "error": repr(ex)[:300]}), flush=True)


if __name__ == "__main__":
    main()
