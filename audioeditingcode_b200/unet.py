"""U-Net evaluation engine: the device-side replacement of what `PipelineWrapper.unet_forward`
(reference code/models.py:160-393, AudioLDM2 variant :691-899) reaches through `self.model.unet.*`.

Host logic (this file) is Python; every FLOP is issued through libaedit.so (ops.CudaOps):
  * convolutions / linears  -> tcgen05 GEMM (implicit conv by TMA where the geometry allows, else patch gather)
  * GroupNorm+SiLU, LayerNorm, GEGLU, attention, resize, layout -> dedicated kernels
Data layout: channels-last.  The residual stream between blocks is fp32; every tensor-core operand is bf16
(produced by the norm / activation kernel in front of each GEMM); accumulation is fp32.

Top-level op order and tap points follow models.py:216-393:
  time embedding -> class embedding (concat) -> conv_in -> down blocks (skip stack) -> mid -> h-space tap /
  replace -> + mid_block_additional_residual -> up blocks (skip replace / zero) -> GN, SiLU, conv_out.
Block internals follow the in-tree statement of the same math (openaimodel.py:175-286 ResBlock,
attention.py:370-469 transformer) — see oracle/unet_torch.py, which this engine is tested against.

Weights arrive under diffusers state-dict names (what the reference's checkpoints contain).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch

from .unet_config import UNetConfig

F32 = torch.float32
NEG_PAD = -1.0e30  # additive bias for key slots that exist only as padding (never present in the reference)


class TextCache:
    """Frozen text conditioning of a run: per cross-attention layer the projected K|V of every text row
    (computed once per prompt set — the reference recomputes to_k/to_v(text) at every step,
    attention.py:234-235), plus the additive key biases (models.py:204-210, :745-755)."""

    def __init__(self, n_rows: int):
        self.n_rows = n_rows
        self.kv: Dict[str, torch.Tensor] = {}       # layer prefix -> bf16 [n_rows, L, 2C]
        # layer prefix -> (KW [R*heads*Lp, C], VW [C, R*heads*Lp], Lp, heads, bias [R, Lp] or None): the cross-attention
        # of the layer folded into two GEMMs (the text is frozen: K.Wq and Wo.V^T are constants of the run)
        self.folded: Dict[str, tuple] = {}
        self.bias: List[Optional[torch.Tensor]] = []  # per stream: fp32 [n_rows, L] or None
        self.lens: List[int] = []


class GraphedForward:
    """One U-Net evaluation of a fixed (B, H, W, text, slots) captured as a CUDA graph: ~10^3 kernel launches
    (and their TMA descriptors, encoded once at capture) replay with a single host call — the launch-latency
    answer SURVEY.md §7 'hard parts' asks for.  Inputs / output live in static buffers."""

    def __init__(self, engine: "UNetEngine", B, H, W, text, slot_map, class_labels, lane: int = 0,
                 capture_stream=None):
        dev = engine.device
        # lane >= 1 = the reverse-process lane (ddm_inversion/inversion_utils._PendingForward): its graph may replay
        # while a forward-process graph is running on another stream, so it owns its workspaces (second CudaOps
        # instance: split-K partial tiles, GroupNorm partials / tickets) and its kernel nodes carry the device's
        # highest launch priority (the sub-wave kernels of the sequential chain take the next free SM slots).
        # lane 2 additionally sizes the GEMM operand rings to fit beside a resident forward-process CTA
        # (ae_set_shared_sm); lane 1 is the variant for the steps that have the machine to themselves.
        self.lane = lane
        with torch.inference_mode(False):   # static buffers outlive the caller's inference_mode scope (copy_ targets)
            self.x = torch.zeros(B, engine.cfg.in_channels, H, W, device=dev, dtype=F32)
            self.t = torch.zeros(B, device=dev, dtype=torch.int64)
            self.slot = None if slot_map is None else torch.empty(slot_map.shape, device=dev, dtype=torch.int32)
            self.cl = None if class_labels is None else torch.empty(class_labels.shape, device=dev, dtype=F32)
        if self.slot is not None:
            self.slot.copy_(slot_map)
        if self.cl is not None:
            self.cl.copy_(class_labels)
        self.text = text
        # Small batches are launch-latency bound (each kernel is a fraction of a wave): with AEDIT_DUAL_STREAM=1 the
        # batch is cut in two halves captured on two forked streams of the same graph, so the two dependency chains
        # interleave on the SMs.  Each chain owns its workspaces (second CudaOps instance).
        self.dual = engine.dual_stream and B % 2 == 0 and B <= engine.dual_stream_max_b and lane == 0
        cur = torch.cuda.current_stream()
        side = capture_stream if capture_stream is not None else torch.cuda.Stream(priority=-1 if lane >= 1 else 0)
        self._side2 = torch.cuda.Stream() if self.dual else None
        main_ops = engine.ops
        # PDL attribute on the norm / attention kernels, per lane (measured on the whole job, profiles/r01_lanes_ab*.log)
        pdlx = int(os.environ.get(("AEDIT_PDLX_FWD", "AEDIT_PDLX_SOLO", "AEDIT_PDLX_SHARED")[lane],
                                  os.environ.get("AEDIT_PDL_EXTRA", str(engine.pdl_extra[lane]))))
        engine.ops.lib.ae_set_pdl_extra(pdlx)
        if lane >= 1:
            engine.ops = engine.ops_b()
            if os.environ.get("AEDIT_REV_PRIORITY", "1") != "0":
                engine.ops.lib.ae_set_launch_priority(engine.ops.lib.ae_greatest_priority())
            engine.ops.lib.ae_set_shared_sm(1 if (lane == 2 and engine.shared_sm_rings) else 0)
        try:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):          # warm-up outside capture: workspaces, smem attributes, lazy allocations
                    self._run(engine)
            cur.wait_stream(side)
            torch.cuda.synchronize()        # also drains the other lane: nothing runs during the capture
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.out = self._run(engine)
        finally:
            if lane >= 1:
                engine.ops.lib.ae_set_launch_priority(0)
                engine.ops.lib.ae_set_shared_sm(0)
            engine.ops = main_ops
            engine.ops.lib.ae_set_pdl_extra(int(os.environ.get("AEDIT_PDL_EXTRA", "0")))

    def _run(self, engine):
        if not self.dual:
            return engine.forward(self.x, self.t, text=self.text, slot_map=self.slot, class_labels=self.cl)
        h = self.x.shape[0] // 2
        sl = (lambda v, a, b: None if v is None else v[a:b])
        cur = torch.cuda.current_stream()
        s2 = self._side2
        out = torch.empty((self.x.shape[0], engine.cfg.out_channels, *self.x.shape[2:]), device=self.x.device, dtype=F32)
        s2.wait_stream(cur)                                            # fork
        engine.forward(self.x[:h], self.t[:h], text=self.text, slot_map=sl(self.slot, 0, h),
                       class_labels=sl(self.cl, 0, h), out=out[:h])
        main_ops = engine.ops
        engine.ops = engine.ops_b()
        try:
            with torch.cuda.stream(s2):
                engine.forward(self.x[h:], self.t[h:], text=self.text, slot_map=sl(self.slot, h, None),
                               class_labels=sl(self.cl, h, None), out=out[h:])
        finally:
            engine.ops = main_ops
        cur.wait_stream(s2)                                            # join
        return out

    def __call__(self, x, t, class_labels=None):
        self.x.copy_(x)
        self.t.copy_(t)
        if class_labels is not None and self.cl is not None:
            self.cl.copy_(class_labels)
        self.graph.replay()
        return self.out


class UNetEngine:
    def __init__(self, cfg: UNetConfig, weights: Dict[str, torch.Tensor], device, ops=None):
        if ops is None:
            from .ops import CudaOps
            ops = CudaOps()
        self.cfg = cfg
        self.ops = ops
        self.device = torch.device(device)
        self.adt = ops.act_dtype            # bf16 on the device path
        self.w: Dict[str, torch.Tensor] = {}
        self._pack(weights)
        # CUDA-graph cache, LRU-bounded: every GraphedForward owns a private memory pool, its static buffers and a reference
        # to its TextCache, so an unbounded cache grows with every new prompt set / clip length of a dataset run
        self._graphs: "OrderedDict" = OrderedDict()
        self.max_graphs = int(os.environ.get("AEDIT_MAX_GRAPHS", "12"))
        self.kernels_per_forward: Dict = {}
        # below this width the 4-D TMA boxes degenerate into 256-byte bursts; gather patches explicitly instead
        self.min_implicit_w = int(os.environ.get("AEDIT_MIN_IMPLICIT_W", "2"))
        self.dual_stream = os.environ.get("AEDIT_DUAL_STREAM", "0") != "0"
        self.dual_stream_max_b = int(os.environ.get("AEDIT_DUAL_STREAM_MAX_B", "4"))
        self._ops_b = None
        self.fold_cross_attn = os.environ.get("AEDIT_FOLD_CROSS_ATTN", "1") != "0" and hasattr(ops, "lib")
        self.graph_placement_tries = int(os.environ.get("AEDIT_GRAPH_PLACEMENT_TRIES", "4"))
        self.shared_sm_rings = os.environ.get("AEDIT_SHARED_SM_RINGS", "1") != "0"
        self.pdl_extra = [0, 0, 15]     # per lane (forward chunks, reverse solo, reverse shared-SM); see GraphedForward
        # GroupNorm statistics from the producing GEMM's epilogue (ae_gemm_args.colstats): every GEMM whose fp32 output
        # feeds a GroupNorm accumulates per-(sample, channel) sums into a slice of one arena that is zeroed once per
        # evaluation; ops.groupnorm then runs a single launch.  AEDIT_GN_COLSTATS=0 restores the statistics kernel.
        self.gn_colstats = os.environ.get("AEDIT_GN_COLSTATS", "1") != "0" and hasattr(ops, "lib")
        self._cs_channels = sum(int(t.shape[0]) for n, t in weights.items()
                                if n.endswith((".conv1.bias", ".conv2.bias", ".proj_out.bias", "conv_in.bias",
                                               ".downsamplers.0.conv.bias", ".upsamplers.0.conv.bias")))
        self._cs_arena = None
        self._cs_off = 0
        self._cs_map: Dict[int, torch.Tensor] = {}
        # probe(name, tensor fp32 [B, H*W, C], (H, W)): optional callback after conv_in and every block of an EAGER
        # forward (error attribution against the oracle's probe of the same name, tools/error_attribution.py)
        self.probe = None

    # ------------------------------------------------------------------------------------------ GroupNorm statistics
    def _cs_begin(self, B: int):
        self._cs_map = {}
        self._cs_off = 0
        self._cs_arena = (torch.zeros(B * self._cs_channels * 2, dtype=torch.int64, device=self.device)
                          if self.gn_colstats else None)

    def _cs_take(self, out: torch.Tensor, B: int, rows_per_sample: int):
        """kwargs for ops.gemm that make it accumulate the column statistics of `out` ([B*rows, N] fp32), or {}."""
        N = out.shape[-1]
        self._cs_map.pop(out.data_ptr(), None)      # a recycled address must not inherit an older tensor's entry
        if self._cs_arena is None or rows_per_sample % 32 != 0 or N % 4 != 0:
            return {}
        n = B * N * 2
        if self._cs_off + n > self._cs_arena.numel():
            return {}
        cs = self._cs_arena[self._cs_off:self._cs_off + n]
        self._cs_off += n
        self._cs_map[out.data_ptr()] = cs
        return dict(colstats=cs, cs_rows=rows_per_sample)

    def _cs_forget(self, *tensors):
        """Tensors that were NOT produced by a statistics-accumulating GEMM (tap / replace paths)."""
        for t in tensors:
            if t is not None:
                self._cs_map.pop(t.data_ptr(), None)

    def _cs_of(self, x: Optional[torch.Tensor]):
        return None if x is None else self._cs_map.get(x.data_ptr())

    def _groupnorm(self, x1, x2, gamma, beta, eps, groups, silu, out, raw_out=None):
        cs1, cs2 = self._cs_of(x1), self._cs_of(x2)
        if cs1 is not None and (x2 is None or cs2 is not None):
            self.ops.groupnorm(x1, x2, gamma, beta, eps, groups, silu, out, raw_out=raw_out, cs1=cs1, cs2=cs2)
        else:
            self.ops.groupnorm(x1, x2, gamma, beta, eps, groups, silu, out, raw_out=raw_out)

    def ops_b(self):
        """Second ops instance (own split-K / GroupNorm workspaces) for the forked chain of a dual-stream graph."""
        if self._ops_b is None:
            self._ops_b = type(self.ops)()
        return self._ops_b

    def graphed(self, B, H, W, text=None, slot_map=None, class_labels=None, slot_key=None, lane: int = 0
                ) -> GraphedForward:
        """Cached CUDA-graph evaluator for this geometry / text binding.  `slot_key`: hashable description of
        slot_map known on the host (avoids a device read-back per call).  `lane`: see GraphedForward."""
        if slot_key is None and slot_map is not None:
            slot_key = tuple(int(v) for v in slot_map.tolist())
        key = (B, H, W, id(text), slot_key, class_labels is not None, lane)
        g = self._graphs.get(key)
        if g is not None:
            self._graphs.move_to_end(key)
        if g is None:
            while len(self._graphs) >= max(1, self.max_graphs):
                self._graphs.popitem(last=False)        # least recently used: frees its pool and unpins its text
            l0 = self.ops.launch_count()
            g = GraphedForward(self, B, H, W, text, slot_map, class_labels, lane=lane)
            g.kernels = (self.ops.launch_count() - l0) // 3       # 2 warm-ups + 1 capture
            if B <= 4 and self.graph_placement_tries > 1:
                g = self._tune_placement(g, (B, H, W, text, slot_map, class_labels), lane)
            self._graphs[key] = g
        return g

    def evict_graphs(self, live_texts=()) -> int:
        """Drop every cached graph whose TextCache is not in `live_texts` (called by the wrappers when they evict text
        conditioning): a graph pins its text K/V and its private pool for as long as it is cached."""
        live = {id(t) for t in live_texts}
        dead = [k for k, g in self._graphs.items() if g.text is not None and id(g.text) not in live]
        for k in dead:
            del self._graphs[k]
        return len(dead)

    def solo_lane(self, B, H, W, text, slot_key, has_cl: bool) -> int:
        """Which reverse-lane graph variant to replay when the lane has the machine to itself: 1 (deep rings, PDL on the
        GEMMs only) unless variant 2 (rings sized for shared SMs, PDL on every family) measured faster ALONE at capture
        time — it does for TANGO-full (6.2 vs 6.8 ms) and AudioLDM-S (2.38 vs 2.46 ms), not for AudioLDM2-large (6.47 vs
        5.97 ms); profiles/r01_other_configs.log.  Decided from the capture-time timings, so it costs nothing per step."""
        g1 = self._graphs.get((B, H, W, id(text), slot_key, has_cl, 1))
        g2 = self._graphs.get((B, H, W, id(text), slot_key, has_cl, 2))
        t1, t2 = getattr(g1, "placement_ms", None), getattr(g2, "placement_ms", None)
        return 2 if (t1 and t2 and min(t2) < 0.99 * min(t1)) else 1

    def _tune_placement(self, g: GraphedForward, args, lane: int) -> GraphedForward:
        """A small-batch evaluation graph (a latency-bound chain of ~900 kernels) runs 4 % slower or faster depending on
        where its private memory pool happens to land (bimodal: 5.97 / 6.20 ms on the same build and box,
        profiles/r01_bimodal_probe.log) — a placement effect of the memory system, not of any setting.  The graph is
        replayed ~100 times per job, so it is captured up to a few times (each capture gets its own pool) and the fastest
        capture is kept."""
        def timed(gr):
            for _ in range(2):
                gr.graph.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(8):
                gr.graph.replay()
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e) / 8
        # every candidate stays alive until the choice is made: a freed pool would simply be handed to the next capture
        cands, times = [g], [timed(g)]
        for i in range(1, self.graph_placement_tries):
            if min(times) < 0.985 * max(times):
                break                                 # both modes seen, the fast one is in hand
            cand = GraphedForward(self, *args, lane=lane)
            cand.kernels = g.kernels
            cands.append(cand)
            times.append(timed(cand))
        best = cands[times.index(min(times))]
        del cands
        best.placement_ms = [round(t, 3) for t in times]
        return best

    # ------------------------------------------------------------------------------------------ weights
    def _to(self, t, dtype):
        return t.detach().to(device=self.device, dtype=dtype).contiguous()

    def _pack(self, w):
        cfg = self.cfg
        P = self.w
        temb_w, temb_b, self.temb_off = [], [], {}
        off = 0
        for name, t in w.items():
            if name.endswith(".time_emb_proj.weight"):
                continue
            if name.endswith(".time_emb_proj.bias"):
                continue
            if name.endswith(".weight") and t.dim() == 4:          # conv: [O,I,kh,kw] -> [O, kh*kw*I]
                P[name] = self._to(t.permute(0, 2, 3, 1).reshape(t.shape[0], -1), self.adt)
            elif name.endswith(".weight") and t.dim() == 2:        # linear
                P[name] = self._to(t, self.adt)
            else:                                                  # biases, norm affine
                P[name] = self._to(t, F32)
        # all ResBlock time-embedding projections as ONE GEMM (K4 of SURVEY.md §2.2)
        for name in sorted(k for k in w if k.endswith(".time_emb_proj.weight")):
            p = name[: -len(".time_emb_proj.weight")]
            temb_w.append(w[name])
            temb_b.append(w[p + ".time_emb_proj.bias"])
            self.temb_off[p] = off
            off += w[name].shape[0]
        self.temb_total = off
        P["__temb_all.weight"] = self._to(torch.cat(temb_w, 0), self.adt)
        P["__temb_all.bias"] = self._to(torch.cat(temb_b, 0), F32)
        # GEGLU feed-forward: interleave value / gate rows in blocks of 16 so the GEMM epilogue (act=2) can form
        # value * gelu(gate) inside one 32-column accumulator chunk (attention.py:37-44 chunk(2) semantics)
        for name in [k for k in w if k.endswith(".ff.net.0.proj.weight")]:
            p = name[: -len(".weight")]
            Wf, bf = w[name], w[p + ".bias"]
            inner = Wf.shape[0] // 2
            assert inner % 16 == 0
            idx = torch.arange(inner).view(-1, 16)
            perm = torch.cat([idx, idx + inner], dim=1).reshape(-1)
            P[p + ".geglu.weight"] = self._to(Wf[perm], self.adt)
            P[p + ".geglu.bias"] = self._to(bf[perm], F32)
        # fused q|k|v (self) and k|v (cross) projections
        for name in [k for k in w if k.endswith(".to_q.weight")]:
            p = name[: -len(".to_q.weight")]
            q, k, v = w[name], w[p + ".to_k.weight"], w[p + ".to_v.weight"]
            if k.shape[1] == q.shape[1] and self._is_self(p):
                P[p + ".qkv.weight"] = self._to(torch.cat([q, k, v], 0), self.adt)
            else:
                P[p + ".kv.weight"] = self._to(torch.cat([k, v], 0), self.adt)

    def _is_self(self, attn_prefix: str) -> bool:
        if attn_prefix.endswith(".attn1"):
            return True
        return self._spec_of(attn_prefix) is None

    def _spec_of(self, attn_prefix: str):
        # ...attentions.{idx}.transformer_blocks.{l}.attn2  -> spec = transformer_specs[idx % n_specs]
        parts = attn_prefix.split(".")
        idx = int(parts[parts.index("attentions") + 1])
        return self.cfg.transformer_specs[idx % len(self.cfg.transformer_specs)]

    # ------------------------------------------------------------------------------------------ text
    def prepare_text(self, streams: Sequence[Optional[torch.Tensor]] = (),
                     masks: Sequence[Optional[torch.Tensor]] = ()) -> TextCache:
        """streams[i]: [R, L_i, D_i] (any float dtype), masks[i]: [R, L_i] (1 keep / 0 discard) or None."""
        cfg = self.cfg
        R = 0
        for s in streams:
            if s is not None:
                R = s.shape[0]
        tc = TextCache(R)
        st_bf = []
        for i, s in enumerate(streams):
            if s is None:
                st_bf.append(None)
                tc.bias.append(None)
                tc.lens.append(0)
                continue
            st_bf.append(self._to(s, self.adt))
            m = masks[i] if i < len(masks) else None
            tc.bias.append(None if m is None else ((1 - m.to(device=self.device, dtype=F32)) * -10000.0).contiguous())
            tc.lens.append(s.shape[1])
        for name in [k for k in self.w if k.endswith(".attn2.kv.weight")]:
            p = name[: -len(".kv.weight")]
            spec = self._spec_of(p)
            s = st_bf[spec[1]]
            Wkv = self.w[name]
            out = self.ops.empty((R, s.shape[1], Wkv.shape[0]), self.adt, self.device)
            self.ops.gemm(s.reshape(-1, s.shape[-1]), Wkv, out_bf16=out.reshape(-1, Wkv.shape[0]))
            tc.kv[p] = out
            if self.fold_cross_attn:
                self._fold_cross_attention(tc, p, out, tc.bias[spec[1]])
        return tc

    def _heads_of(self, prefix: str) -> int:
        parts = prefix.split(".")
        nlev = len(self.cfg.block_out_channels)
        if parts[0] == "down_blocks":
            level = int(parts[1])
        elif parts[0] == "up_blocks":
            level = nlev - 1 - int(parts[1])
        else:
            level = nlev - 1
        return self.cfg.num_heads[level]

    def _fold_cross_attention(self, tc: TextCache, p: str, kv: torch.Tensor, bias: Optional[torch.Tensor]):
        """Cross-attention against frozen text as two GEMMs (include/aedit.h, ae_gemm_args.sm_*): precompute
        KW[(r,h,l), c] = scale * sum_j K_r[l, h*d+j] Wq[h*d+j, c] and VW[c, (r,h,l)] = sum_j Wo[c, h*d+j] V_r[l, h*d+j]
        with the library's own GEMM (one batched call over the heads per text row)."""
        ops = self.ops
        R, L, C2 = kv.shape
        C = C2 // 2
        heads = self._heads_of(p)
        d = C // heads
        Lp = 8 if L <= 8 else (16 if L <= 16 else (32 if L <= 32 else 0))
        if Lp == 0 or (d * 2) % 16 != 0 or (R * heads * Lp) % 32 != 0:
            return
        WqT = self.w.get(p + ".to_q.weightT")
        if WqT is None:
            WqT = self.w[p + ".to_q.weight"].t().contiguous()        # [C_in, C_out]: row c, column h*d+j
            self.w[p + ".to_q.weightT"] = WqT
        Wo = self.w[p + ".to_out.0.weight"]                             # [C_out, C_in = heads*d]
        NK = R * heads * Lp
        KW = torch.zeros((NK, C), dtype=self.adt, device=self.device)
        VW = torch.zeros((C, NK), dtype=self.adt, device=self.device)
        scale = float(d) ** -0.5
        for r in range(R):
            # KW rows of text row r: per head z, A = K_r[:, z*d:(z+1)*d] (L x d), W = WqT[:, z*d:(z+1)*d] (C x d)
            ops.gemm(kv[r], WqT, out_bf16=KW[r * heads * Lp:], M=L, K=d, lda=C2, ldw=C, batch=heads, strideA=d,
                     strideW=d, stride_out=Lp * C, ld_out_bf16=C, alpha=scale, w_dynamic=True)
            # VW columns of text row r: per head z, A = Wo[:, z*d:(z+1)*d] (C x d), W = V_r[:, z*d:(z+1)*d] (L x d)
            ops.gemm(Wo, kv[r, :, C:], out_bf16=VW[:, r * heads * Lp:], M=C, K=d, lda=C, ldw=C2, batch=heads,
                     strideA=d, strideW=d, stride_out=Lp, ld_out_bf16=NK, w_dynamic=True)
        bias_p = None
        if bias is not None or L != Lp:
            bias_p = torch.full((R, Lp), NEG_PAD, dtype=F32, device=self.device)
            bias_p[:, :L] = 0.0 if bias is None else bias.to(F32)
        tc.folded[p] = (KW, VW, Lp, heads, bias_p)

    # ------------------------------------------------------------------------------------------ blocks
    def _conv3x3(self, a_bf16, B, H, W, Cin, name, out, rowbias=None, rows_per_group=1, residual=None, stats=False):
        """a_bf16: [B,H,W,Cin] channels-last operand; out fp32 [B*H*W, Cout].  stats: the output feeds a GroupNorm."""
        ops = self.ops
        Wt, bias = self.w[name + ".weight"], self.w[name + ".bias"]
        cs = self._cs_take(out, B, H * W) if stats else {}
        if W >= self.min_implicit_w and ops.conv_supported(B, H, W, Cin):
            ops.gemm(a_bf16, Wt, out_f32=out, bias=bias, rowbias=rowbias, rows_per_group=rows_per_group,
                     residual=residual, conv=(B, H, W, Cin, 3, 3, 1, 1), **cs)
        else:
            K = 9 * Cin
            ld = (K + 7) // 8 * 8
            col = ops.empty((B * H * W, ld), self.adt, self.device)
            ops.im2col(a_bf16, B, H, W, Cin, 3, 3, 1, 1, 1, 1, H, W, col)
            ops.gemm(col, Wt, out_f32=out, bias=bias, rowbias=rowbias, rows_per_group=rows_per_group,
                     residual=residual, K=K, **cs)

    def _resnet(self, x1, x2, B, H, W, p, temb_all):
        ops, cfg = self.ops, self.cfg
        C1 = x1.shape[-1]
        C2 = 0 if x2 is None else x2.shape[-1]
        Cin = C1 + C2
        Cout = self.w[p + ".conv1.bias"].shape[0]
        M = B * H * W
        has_sc = (p + ".conv_shortcut.weight") in self.w
        a1 = ops.empty((B, H, W, Cin), self.adt, self.device)
        raw = ops.empty((M, Cin), self.adt, self.device) if has_sc else None
        self._groupnorm(x1, x2, self.w[p + ".norm1.weight"], self.w[p + ".norm1.bias"], cfg.norm_eps,
                        cfg.norm_num_groups, True, a1, raw_out=raw)
        h = ops.empty((M, Cout), F32, self.device)
        off = self.temb_off[p]
        self._conv3x3(a1, B, H, W, Cin, p + ".conv1", h, rowbias=temb_all[:, off:off + Cout], rows_per_group=H * W,
                      stats=True)
        a2 = ops.empty((B, H, W, Cout), self.adt, self.device)
        self._groupnorm(h.view(B, H * W, Cout), None, self.w[p + ".norm2.weight"], self.w[p + ".norm2.bias"],
                        cfg.norm_eps, cfg.norm_num_groups, True, a2)
        if has_sc:
            res = ops.empty((M, Cout), F32, self.device)
            ops.gemm(raw, self.w[p + ".conv_shortcut.weight"], out_f32=res, bias=self.w[p + ".conv_shortcut.bias"])
        else:
            res = x1.reshape(M, Cout)
        out = ops.empty((M, Cout), F32, self.device)
        self._conv3x3(a2, B, H, W, Cout, p + ".conv2", out, residual=res, stats=True)
        return out.view(B, H * W, Cout)

    def _attention(self, q, k, v, out, heads, B, Tq, Tk, ld_q, bs_q, ld_k, bs_k, ld_v, bs_v, kv_map=None, bias=None):
        d = out.shape[-1] // heads
        self.ops.attention(q, k, v, out, heads, d, float(d) ** -0.5, Tq, Tk, B, ld_q, bs_q, ld_k, bs_k, ld_v, bs_v,
                           kv_map=kv_map, bias=bias)

    def _transformer(self, x, B, H, W, p, heads, spec, text: Optional[TextCache], slot_map):
        ops, cfg = self.ops, self.cfg
        C = x.shape[-1]
        T = H * W
        M = B * T
        g = ops.empty((M, C), self.adt, self.device)
        self._groupnorm(x, None, self.w[p + ".norm.weight"], self.w[p + ".norm.bias"], 1e-6, cfg.norm_num_groups, False, g)
        hs = ops.empty((M, C), F32, self.device)
        ops.gemm(g, self.w[p + ".proj_in.weight"], out_f32=hs, bias=self.w[p + ".proj_in.bias"])
        hs_b = None
        nl = cfg.transformer_layers_per_block
        for l in range(nl):
            q = f"{p}.transformer_blocks.{l}"
            # --- attn1 (self)
            n = ops.empty((M, C), self.adt, self.device)
            ops.layernorm(hs, self.w[q + ".norm1.weight"], self.w[q + ".norm1.bias"], n)
            qkv = ops.empty((M, 3 * C), self.adt, self.device)
            ops.gemm(n, self.w[q + ".attn1.qkv.weight"], out_bf16=qkv)
            a = ops.empty((M, C), self.adt, self.device)
            self._attention(qkv, qkv[:, C:], qkv[:, 2 * C:], a, heads, B, T, T, 3 * C, T * 3 * C, 3 * C, T * 3 * C,
                            3 * C, T * 3 * C)
            ops.gemm(a, self.w[q + ".attn1.to_out.0.weight"], out_f32=hs, bias=self.w[q + ".attn1.to_out.0.bias"],
                     residual=hs)
            # --- attn2 (self or cross)
            folded = False
            ops.layernorm(hs, self.w[q + ".norm2.weight"], self.w[q + ".norm2.bias"], n)
            if spec is None:
                ops.gemm(n, self.w[q + ".attn2.qkv.weight"], out_bf16=qkv)
                self._attention(qkv, qkv[:, C:], qkv[:, 2 * C:], a, heads, B, T, T, 3 * C, T * 3 * C, 3 * C,
                                T * 3 * C, 3 * C, T * 3 * C)
            elif text is not None and (q + ".attn2") in text.folded:
                # frozen text: scores = n . KW^T with the per-head softmax in the epilogue, out = P . VW^T (+ bias + hs)
                KW, VW, Lp, fh, fbias = text.folded[q + ".attn2"]
                sm = slot_map if slot_map is not None else self._identity_slots(B)
                P = ops.empty((M, KW.shape[0]), self.adt, self.device)
                ops.gemm(n, KW, out_bf16=P, softmax=(Lp, fh * Lp, sm, T, fbias))
                ops.gemm(P, VW, out_f32=hs, bias=self.w[q + ".attn2.to_out.0.bias"], residual=hs)
                folded = True
            else:
                if text is None or (q + ".attn2") not in text.kv:
                    raise ValueError("cross-attention layer needs prepared text (UNetEngine.prepare_text)")
                qq = qkv[:, :C]
                ops.gemm(n, self.w[q + ".attn2.to_q.weight"], out_bf16=qq)
                kv = text.kv[q + ".attn2"]
                L = kv.shape[1]
                self._attention(qq, kv, kv[:, :, C:], a, heads, B, T, L, 3 * C, T * 3 * C, 2 * C, L * 2 * C, 2 * C,
                                L * 2 * C, kv_map=slot_map, bias=text.bias[spec[1]])
            if not folded:
                ops.gemm(a, self.w[q + ".attn2.to_out.0.weight"], out_f32=hs, bias=self.w[q + ".attn2.to_out.0.bias"],
                         residual=hs)
            # --- GEGLU feed-forward
            ops.layernorm(hs, self.w[q + ".norm3.weight"], self.w[q + ".norm3.bias"], n)
            gg = ops.empty((M, 4 * C), self.adt, self.device)
            ops.gemm(n, self.w[q + ".ff.net.0.proj.geglu.weight"], out_bf16=gg,
                     bias=self.w[q + ".ff.net.0.proj.geglu.bias"], act=2)
            if l == nl - 1:
                hs_b = ops.empty((M, C), self.adt, self.device)
            ops.gemm(gg, self.w[q + ".ff.net.2.weight"], out_f32=hs, out_bf16=hs_b if l == nl - 1 else None,
                     bias=self.w[q + ".ff.net.2.bias"], residual=hs)
        out = ops.empty((M, C), F32, self.device)
        ops.gemm(hs_b, self.w[p + ".proj_out.weight"], out_f32=out, bias=self.w[p + ".proj_out.bias"],
                 residual=x.reshape(M, C), **self._cs_take(out, B, T))
        return out.view(B, T, C)

    def _identity_slots(self, B: int) -> torch.Tensor:
        t = self.__dict__.setdefault("_id_slots", {}).get(B)
        if t is None:
            t = torch.arange(B, dtype=torch.int32, device=self.device)
            self._id_slots[B] = t
        return t

    def _site(self, x, B, H, W, base, idx0, level, text, slot_map):
        cfg = self.cfg
        ns = len(cfg.transformer_specs)
        for j, spec in enumerate(cfg.transformer_specs):
            x = self._transformer(x, B, H, W, f"{base}.{idx0 * ns + j}", cfg.num_heads[level], spec, text, slot_map)
        return x

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, sample: torch.Tensor, timesteps: torch.Tensor, text: Optional[TextCache] = None,
                slot_map: Optional[torch.Tensor] = None, class_labels: Optional[torch.Tensor] = None,
                mid_block_additional_residual=None, replace_h_space=None, replace_skip_conns=None,
                zero_out_resconns=None, want_taps: bool = False, out: Optional[torch.Tensor] = None):
        """sample: fp32 NCHW [B,Cin,H,W]; timesteps: int64 [B]; slot_map: int32 [B] rows of `text`;
        class_labels: [B, class_embed_dim].  Returns eps (fp32 NCHW) and, if want_taps, (eps, h_space, skips)
        in the reference's NCHW convention (models.py:336-361,393)."""
        ops, cfg = self.ops, self.cfg
        dev = self.device
        sample = sample.to(dev, F32).contiguous()
        B, Cin, H, W = sample.shape
        timesteps = torch.as_tensor(timesteps).to(dev, torch.int64).reshape(-1)
        if timesteps.numel() == 1 and B > 1:
            timesteps = timesteps.expand(B)
        timesteps = timesteps.contiguous()
        if slot_map is not None:
            slot_map = slot_map.to(dev, torch.int32).contiguous()
        ch = cfg.block_out_channels
        nlev = len(ch)
        ted = 4 * ch[0]
        temb_ch = self.w["__temb_all.weight"].shape[1]
        self._cs_begin(B)

        # ---- time / class embedding (models.py:231-256); silu(emb) is the only consumer of emb
        tproj = ops.empty((B, ch[0]), self.adt, dev)
        ops.timestep_embedding(timesteps, ch[0], tproj)
        e1 = ops.empty((B, ted), self.adt, dev)
        ops.gemm(tproj, self.w["time_embedding.linear_1.weight"], out_bf16=e1,
                 bias=self.w["time_embedding.linear_1.bias"], act=1)
        emb_act = ops.empty((B, temb_ch), self.adt, dev)
        ops.gemm(e1, self.w["time_embedding.linear_2.weight"], out_bf16=emb_act[:, :ted],
                 bias=self.w["time_embedding.linear_2.bias"], act=1)
        if cfg.class_embed_dim is not None:
            if class_labels is None:
                raise ValueError("class_labels should be provided when num_class_embeds > 0")  # models.py:241-242
            cl = ops.empty((B, cfg.class_embed_dim), self.adt, dev)
            ops.cast_bf16(class_labels.to(dev, F32).contiguous(), cl)
            if not cfg.class_embeddings_concat:
                raise NotImplementedError("additive class embedding is not used by the audio models")
            ops.gemm(cl, self.w["class_embedding.weight"], out_bf16=emb_act[:, ted:],
                     bias=self.w["class_embedding.bias"], act=1)
        temb_all = ops.empty((B, self.temb_total), F32, dev)
        ops.gemm(emb_act, self.w["__temb_all.weight"], out_f32=temb_all, bias=self.w["__temb_all.bias"])

        # ---- conv_in (Cin = 8: explicit patch gather, K = 72)
        x_nhwc = ops.empty((B, H, W, Cin), F32, dev)
        ops.nchw_to_nhwc(sample, out_f32=x_nhwc)
        K0 = 9 * Cin
        col = ops.empty((B * H * W, (K0 + 7) // 8 * 8), self.adt, dev)
        ops.im2col(x_nhwc, B, H, W, Cin, 3, 3, 1, 1, 1, 1, H, W, col)
        h = ops.empty((B * H * W, ch[0]), F32, dev)
        ops.gemm(col, self.w["conv_in.weight"], out_f32=h, bias=self.w["conv_in.bias"], K=K0,
                 **self._cs_take(h, B, H * W))
        h = h.view(B, H * W, ch[0])
        probe = self.probe if self.probe is not None else (lambda name, t, hw: None)
        probe("conv_in", h, (H, W))

        sizes = [(H, W)]
        skips = [h]
        hh, ww = H, W
        for i in range(nlev):
            for j in range(cfg.layers_per_block):
                h = self._resnet(h, None, B, hh, ww, f"down_blocks.{i}.resnets.{j}", temb_all)
                probe(f"down_blocks.{i}.resnets.{j}", h, (hh, ww))
                if cfg.attn_levels[i]:
                    h = self._site(h, B, hh, ww, f"down_blocks.{i}.attentions", j, i, text, slot_map)
                    probe(f"down_blocks.{i}.attentions.{j}", h, (hh, ww))
                skips.append(h)
            if i != nlev - 1:
                C = ch[i]
                ho, wo = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
                col = ops.empty((B * ho * wo, 9 * C), self.adt, dev)
                ops.im2col(h, B, hh, ww, C, 3, 3, 2, 1, 1, 1, ho, wo, col)
                d = ops.empty((B * ho * wo, C), F32, dev)
                p = f"down_blocks.{i}.downsamplers.0.conv"
                ops.gemm(col, self.w[p + ".weight"], out_f32=d, bias=self.w[p + ".bias"], **self._cs_take(d, B, ho * wo))
                hh, ww = ho, wo
                sizes.append((hh, ww))
                h = d.view(B, hh * ww, C)
                probe(f"down_blocks.{i}.downsamplers.0", h, (hh, ww))
                skips.append(h)

        h = self._resnet(h, None, B, hh, ww, "mid_block.resnets.0", temb_all)
        probe("mid_block.resnets.0", h, (hh, ww))
        h = self._site(h, B, hh, ww, "mid_block.attentions", 0, nlev - 1, text, slot_map)
        probe("mid_block.attentions.0", h, (hh, ww))
        h = self._resnet(h, None, B, hh, ww, "mid_block.resnets.1", temb_all)
        probe("mid_block.resnets.1", h, (hh, ww))

        # ---- h-space tap / replace, additive residual (models.py:336-343)
        Cm = ch[-1]
        h_space = None
        if replace_h_space is not None:
            h_space = replace_h_space
            rep = ops.empty((B, hh, ww, Cm), F32, dev)
            ops.nchw_to_nhwc(replace_h_space.to(dev, F32).expand(B, -1, -1, -1).contiguous(), out_f32=rep)
            h = rep.view(B, hh * ww, Cm)
            self._cs_forget(h)
        elif want_taps:
            h_space = ops.empty((B, Cm, hh, ww), F32, dev)
            ops.nhwc_to_nchw(h, B, Cm, hh, ww, h_space)
        if mid_block_additional_residual is not None:
            add = ops.empty((B, hh, ww, Cm), F32, dev)
            ops.nchw_to_nhwc(mid_block_additional_residual.to(dev, F32).expand(B, -1, -1, -1).contiguous(), out_f32=add)
            h2 = ops.empty((B, hh * ww, Cm), F32, dev)
            ops.add(h, add.view(B, hh * ww, Cm), h2)
            h = h2
            self._cs_forget(h)

        extracted = {}
        n_up = cfg.layers_per_block + 1
        for i in range(nlev):
            level = nlev - 1 - i
            hh, ww = sizes[level]
            res = skips[-n_up:]
            skips = skips[:-n_up]
            if replace_skip_conns is not None and replace_skip_conns.get(i):
                res = [self._from_nchw(t, B) for t in replace_skip_conns.get(i)]
                self._cs_forget(*res)
            if zero_out_resconns is not None:
                if (type(zero_out_resconns) is int and i >= (zero_out_resconns - 1)) or \
                        (type(zero_out_resconns) is list and i in zero_out_resconns):
                    res = [torch.zeros_like(t) for t in res]
                    self._cs_forget(*res)
            if want_taps:
                extracted[i] = [self._to_nchw(t, B, hh, ww) for t in res]
            res = list(res)
            for j in range(n_up):
                h = self._resnet(h, res.pop(), B, hh, ww, f"up_blocks.{i}.resnets.{j}", temb_all)
                probe(f"up_blocks.{i}.resnets.{j}", h, (hh, ww))
                if cfg.attn_levels[level]:
                    h = self._site(h, B, hh, ww, f"up_blocks.{i}.attentions", j, level, text, slot_map)
                    probe(f"up_blocks.{i}.attentions.{j}", h, (hh, ww))
            if i != nlev - 1:
                C = ch[level]
                ho, wo = sizes[level - 1]
                up = ops.empty((B, ho, wo, C), self.adt, dev)
                ops.upsample_nearest(h, B, hh, ww, C, ho, wo, up)
                u = ops.empty((B * ho * wo, C), F32, dev)
                self._conv3x3(up, B, ho, wo, C, f"up_blocks.{i}.upsamplers.0.conv", u, stats=True)
                h = u.view(B, ho * wo, C)
                probe(f"up_blocks.{i}.upsamplers.0", h, (ho, wo))

        # ---- conv_norm_out -> SiLU -> conv_out (models.py:385-388)
        a = ops.empty((B, H, W, ch[0]), self.adt, dev)
        self._groupnorm(h, None, self.w["conv_norm_out.weight"], self.w["conv_norm_out.bias"], cfg.norm_eps,
                        cfg.norm_num_groups, True, a)
        Co = cfg.out_channels
        o = ops.empty((B * H * W, Co), F32, dev)
        self._conv3x3(a, B, H, W, ch[0], "conv_out", o)
        if out is None:
            out = ops.empty((B, Co, H, W), F32, dev)
        ops.nhwc_to_nchw(o, B, Co, H, W, out)
        if want_taps:
            return out, h_space, extracted
        return out

    def _to_nchw(self, t, B, H, W):
        C = t.shape[-1]
        o = self.ops.empty((B, C, H, W), F32, self.device)
        self.ops.nhwc_to_nchw(t.contiguous(), B, C, H, W, o)
        return o

    def _from_nchw(self, t, B):
        t = t.to(self.device, F32)
        if t.shape[0] != B:
            t = t.expand(B, -1, -1, -1)
        t = t.contiguous()
        _, C, H, W = t.shape
        o = self.ops.empty((B, H, W, C), F32, self.device)
        self.ops.nchw_to_nhwc(t, out_f32=o)
        return o.view(B, H * W, C)
