"""CPU-only tests of the host side: C-ABI surface, scheduler table (integer index math + fp32 scalars, bit-exact
against the reference's own get_variance / get_alpha_prod_t_prev golden values), FLOP accounting, weight
inventory, model-id dispatch."""
import ctypes as C
import os
import re

import pytest
import torch

from tests.helpers import load_golden
from audioeditingcode_b200 import _lib, unet_config as UC, weights as W
from audioeditingcode_b200.flops import count_flops
from audioeditingcode_b200.scheduler import DDIMScheduler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "aedit.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ae_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libaedit.so does not export {name}"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert lib.ae_version() >= 100


@pytest.mark.parametrize("n", [50, 100, 200])
def test_sched_table_bitexact_vs_reference_scalars(n):
    g = load_golden(f"sched_{n}.npz")
    s = DDIMScheduler()
    s.set_timesteps(n)
    assert torch.equal(s.timesteps, g["timesteps"])
    assert torch.equal(s.alphas_cumprod, g["alphas_cumprod"])
    tab = s.table
    for pos in range(n):
        r = tab.row(pos)
        t = int(g["timesteps"][pos])
        assert r.t == t and r.prev_t == t - 1000 // n                 # models.py:96-97 integer math
        assert tab.pos_of_t(t) == pos                                 # t_to_idx, inversion_utils.py:68
        assert r.variance == float(g["variance"][pos])               # models.py:539-545, fp32 op order
        assert r.alpha_prod_t_prev == float(g["alpha_prod_t_prev"][pos])   # models.py:547-549 (final_alpha at prev<0)
        ab = float(g["alphas_cumprod"][t])
        assert r.alpha_bar_t == ab
        assert r.sqrt_ab == float(torch.tensor(ab) ** 0.5)
        assert r.sqrt_1mab == float((1 - torch.tensor(ab)) ** 0.5)
        assert r.sqrt_var == float(torch.tensor(r.variance) ** 0.5)
    assert tab.lib.ae_sched_pos_of_t(tab.h, 2) == -1
    with pytest.raises(KeyError):
        tab.pos_of_t(2)


def test_flop_accounting_matches_survey():
    f = count_flops(UC.preset("audioldm-s"), 256, 16, 1, ())
    assert round(f["conv"] / 1e9, 1) == 64.2 and round(f["linear"] / 1e9, 1) == 27.3 and round(f["attn"] / 1e9, 1) == 11.9
    assert round(f["total"] / 1e9, 1) == 103.4


def test_weight_inventory_and_param_counts():
    from oracle import unet_torch as U
    for name, millions in [("audioldm-s", 185.0), ("audioldm-l", 739.1), ("audioldm2", 346.9), ("tango", 865.9)]:
        cfg = UC.preset(name)
        assert W.weight_shapes(cfg) == U.weight_shapes(cfg)
        assert round(W.count_params(cfg) / 1e6, 1) == millions
    cfg = UC.preset("tiny-audioldm2")
    a, b = W.synthetic_weights(cfg, 0), U.synthetic_weights(cfg, 0)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


def test_model_id_dispatch_and_errors():
    assert UC.from_model_id("cvssp/audioldm-s-full-v2").name == "audioldm-s"
    assert UC.from_model_id("cvssp/audioldm-l-full").name == "audioldm-l"
    assert UC.from_model_id("cvssp/audioldm2-large").name == "audioldm2-large"
    assert UC.from_model_id("cvssp/audioldm2-music").name == "audioldm2"
    assert UC.from_model_id("declare-lab/tango-full-ft-audiocaps").prediction_type == "v_prediction"
    with pytest.raises(ValueError):
        UC.from_model_id("CompVis/stable-diffusion-v1-4")


def test_conv_fast_path_geometry():
    lib = _lib.load()
    ok = lambda B, H, W_, C_: bool(lib.ae_gemm_conv_supported(B, H, W_, C_))
    assert ok(2, 256, 16, 128) and ok(2, 128, 8, 256) and ok(2, 64, 4, 384) and ok(2, 32, 2, 640)   # AudioLDM-S levels
    assert ok(1, 1, 1024, 64) and ok(1, 1024, 64, 128)                                             # vocoder / VAE
    assert not ok(2, 256, 16, 8) and not ok(1, 141, 16, 128) and not ok(1, 20, 16, 128)


def test_forward_chunk_plan():
    """inversion_utils._chunk_plan: a partition of the N loop positions; with a hint, row `hint` tops the first
    chunk, the rows below follow in descending order, the rows above come last."""
    from audioeditingcode_b200.ddm_inversion.inversion_utils import _chunk_plan
    for N, tb in [(200, 50), (200, 25), (50, 50), (12, 5), (7, 3), (10, 1), (13, 4)]:
        natural = [(p, min(tb, N - p)) for p in range(0, N, tb)]
        assert _chunk_plan(N, tb, None) == natural
        for hint in range(0, N + 1):
            plan = _chunk_plan(N, tb, hint)
            pos = sorted(q for p, c in plan for q in range(p, p + c))
            assert pos == list(range(N)), (N, tb, hint, plan)
            if tb == 1:
                assert plan == natural
                continue
            rows = [(N - p - c, N - p - 1) for p, c in plan]          # (lo, hi) in idx space
            top = min(hint, N - 1)
            assert rows[0][1] == top
            below = [r for r in rows if r[1] <= top]
            assert rows[:len(below)] == below and below == sorted(below, key=lambda r: -r[1])
            assert all(c <= tb + max(1, tb // 4) for _, c in plan)
    assert _chunk_plan(200, 50, 100, head=0) == [(99, 50), (149, 51), (49, 50), (0, 49)]
    assert _chunk_plan(200, 50, 100, head=10) == [(99, 10), (109, 50), (159, 41), (49, 50), (0, 49)]
    assert _chunk_plan(200, 50, 200, head=0) == [(0, 50), (50, 50), (100, 50), (150, 50)]
    assert _chunk_plan(200, 50, 200) == [(0, 10), (10, 50), (60, 50), (110, 50), (160, 40)]


def test_pending_forward_guard():
    """The overlap fast path is refused unless the reverse process receives the forward process's own, unmodified
    tensors (storage + version counters) and the same eta table."""
    import torch
    from audioeditingcode_b200.ddm_inversion.inversion_utils import _PendingForward
    with torch.inference_mode():
        with torch.inference_mode(False):
            zs, xts = torch.zeros(4, 3), torch.zeros(5, 3)
        p = _PendingForward()
        p.zs_ptr, p.zs_ver, p.xts_ptr, p.xts_ver, p.eta_key, p.N = zs.data_ptr(), zs._version, xts.data_ptr(), xts._version, (1.0,) * 4, 4
        p.chunks = [(2, 3, "e1"), (0, 1, "e0")]
        assert p.matches(zs[:2], xts, (1.0,) * 4)
        assert not p.matches(zs[1:3], xts, (1.0,) * 4)            # different rows
        assert not p.matches(zs[:2], xts, (0.5,) * 4)             # different eta table
        assert not p.matches(zs[:2].clone(), xts, (1.0,) * 4)     # a copy
        assert not p.matches(torch.zeros(4, 3), xts, (1.0,) * 4)  # inference tensor: no version counter
        assert p.event_for(3) == "e1" and p.event_for(0) == "e0" and p.event_for(9) is None
        zs[1] += 1
        assert not p.matches(zs[:2], xts, (1.0,) * 4)             # modified in place after the forward process


def test_solo_lane_choice_from_capture_timings():
    """UNetEngine.solo_lane: variant 2 only when both variants were timed at capture and 2 was faster alone."""
    from audioeditingcode_b200.unet import UNetEngine

    class G:
        def __init__(self, ms):
            if ms is not None:
                self.placement_ms = ms
    eng = UNetEngine.__new__(UNetEngine)
    text = object()
    key = lambda lane: (2, 256, 16, id(text), ("rev", 1), False, lane)
    eng._graphs = {}
    assert eng.solo_lane(2, 256, 16, text, ("rev", 1), False) == 1            # nothing captured yet
    eng._graphs[key(2)] = G([6.2, 6.19])
    assert eng.solo_lane(2, 256, 16, text, ("rev", 1), False) == 1            # variant 1 not timed yet
    eng._graphs[key(1)] = G([6.8, 6.81])
    assert eng.solo_lane(2, 256, 16, text, ("rev", 1), False) == 2            # TANGO-like: the shared variant wins alone
    eng._graphs[key(1)] = G([5.97, 6.2])
    eng._graphs[key(2)] = G([6.47])
    assert eng.solo_lane(2, 256, 16, text, ("rev", 1), False) == 1            # AudioLDM2-large-like
    eng._graphs[key(1)] = G(None)
    assert eng.solo_lane(2, 256, 16, text, ("rev", 1), False) == 1            # untimed graph (placement tuning off)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys,
    timed on the oracle port, no GPU needed.  Tiny workload so the whole CPU suite stays fast."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "tiny",
                        "--steps", "1", "--warmup", "1", "--cpu-steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "denoising-steps/sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"] == "tiny" and "model" not in line["config"]


def test_c_abi_argument_errors_without_a_device():
    """Error behaviour of the C ABI (include/aedit.h: negative AE_E* + ae_last_error message).  Every call here is
    rejected by argument validation BEFORE the library touches the device, so it runs without a GPU (the pointers are
    dummies that are never dereferenced on the host)."""
    import ctypes as C
    lib = _lib.load()
    dummy = 0x10000
    a = _lib.AeGemmArgs()
    assert lib.ae_gemm(C.byref(a), None) == -1 and b"null operand" in lib.ae_last_error()
    a.A, a.W, a.M, a.N, a.K, a.lda, a.ldw = dummy, dummy, 128, 128, 64, 64, 64
    assert lib.ae_gemm(C.byref(a), None) == -1 and b"no output" in lib.ae_last_error()
    a.out_bf16 = dummy
    a.act = 4
    assert lib.ae_gemm(C.byref(a), None) == -1 and b"act must be" in lib.ae_last_error()
    a.act = 3                                   # grouped softmax without its geometry
    assert lib.ae_gemm(C.byref(a), None) == -1 and b"grouped-softmax" in lib.ae_last_error()
    a.sm_L, a.sm_block, a.sm_rows, a.sm_slot = 8, 64, 64, dummy
    a.out_f32 = dummy                           # act 3 takes a plain bf16 output
    assert lib.ae_gemm(C.byref(a), None) == -1 and b"plain bf16 output" in lib.ae_last_error()
    a.act, a.sm_L, a.sm_block, a.sm_rows, a.sm_slot = 0, 0, 0, 0, None
    a.ld_out_f32 = 128
    a.colstats, a.cs_rows_per_sample = dummy, 48     # rows per sample must be a multiple of 32 dividing M
    assert lib.ae_gemm(C.byref(a), None) == -1 and b"multiple of 32" in lib.ae_last_error()
    a.colstats, a.cs_rows_per_sample = None, 0
    a.act, a.N = 2, 100                         # GEGLU needs N % 32 == 0
    assert lib.ae_gemm(C.byref(a), None) == -1 and b"GEGLU" in lib.ae_last_error()
    # attention: a head dim that is not instantiated, and misaligned strides
    rc = lib.ae_attention(dummy, 64, 64, dummy, 64, 64, dummy, 64, 64, None, None, 0, 1, 1, 50, 64, 64, 1.0, dummy, 64, 64,
                          None)
    assert rc == -3 and b"head dim 50" in lib.ae_last_error()
    rc = lib.ae_attention(dummy, 60, 64, dummy, 64, 64, dummy, 64, 64, None, None, 0, 1, 1, 64, 64, 64, 1.0, dummy, 64, 64,
                          None)
    assert rc == -1 and b"alignment" in lib.ae_last_error()
    # GroupNorm from column statistics without the statistics
    rc = lib.ae_groupnorm_cs(dummy, 64, None, None, 0, None, 1, 64, 32, 1e-5, dummy, dummy, 1, dummy, None, None, dummy, None)
    assert rc == -1 and b"column statistics missing" in lib.ae_last_error()
    rc = lib.ae_groupnorm(dummy, 60, None, 0, 1, 64, 32, 1e-5, dummy, dummy, 1, dummy, None, None, dummy, None)
    assert rc == -1 and b"divisible" in lib.ae_last_error()
    assert lib.ae_layernorm(dummy, 4, 4096, 1e-5, dummy, dummy, dummy, None) == -1


def test_pc_drift_finite_difference_step_resolution(monkeypatch):
    """get_eigenvectors' step: argument > AEDIT_PC_FD_CONST > the evaluator's pc_fd_const > the caller's const (the
    reference's behaviour, pc_drift.py:130,140); the wrappers' own step follows the operand type of the loaded library."""
    import types
    from audioeditingcode_b200 import pc_drift as PC, models, _lib
    plain, tuned = types.SimpleNamespace(), types.SimpleNamespace(pc_fd_const=2.0)
    monkeypatch.delenv("AEDIT_PC_FD_CONST", raising=False)
    assert PC.resolve_fd_const(plain, 1e-3) == 1e-3
    assert PC.resolve_fd_const(tuned, 1e-3) == 2.0
    assert PC.resolve_fd_const(tuned, 1e-3, 0.5) == 0.5
    monkeypatch.setenv("AEDIT_PC_FD_CONST", "reference")
    assert PC.resolve_fd_const(tuned, 1e-3) == 1e-3
    monkeypatch.setenv("AEDIT_PC_FD_CONST", "4")
    assert PC.resolve_fd_const(tuned, 1e-3) == 4.0 and PC.resolve_fd_const(tuned, 1e-3, 0.5) == 0.5
    monkeypatch.delenv("AEDIT_PC_FD_CONST")
    want = 8.0 if _lib.load().ae_operand_dtype() == 0 else 1.0
    w = models.PipelineWrapper.__new__(models.PipelineWrapper)      # the property needs no loaded model
    torch.nn.Module.__init__(w)
    assert w.pc_fd_const == want
    w.pc_fd_const = None
    assert w.pc_fd_const is None and PC.resolve_fd_const(w, 1e-3) == 1e-3
