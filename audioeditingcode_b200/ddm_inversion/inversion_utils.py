"""Edit-friendly DDPM inversion loops — drop-in for code/ddm_inversion/inversion_utils.py (same names,
signatures, return tuples, error behaviour):

    inversion_forward_process(model, x0, etas, prog_bar, prompts, cfg_scales, num_inference_steps, cutoff_points,
                              numerical_fix, extract_h_space, extract_skipconns, duration, first_order)
        -> (xt, zs, xts, extra_info[, hspaces[, skipconns]])                     reference :8-144
    inversion_reverse_process(model, xT, tstart, fix_alpha, etas, prompts, neg_prompts, cfg_scales, prog_bar, zs,
                              cutoff_points, hspace_add, hspace_replace, skipconns_replace, zero_out_resconns,
                              extract_h_space, extract_skipconns, duration, first_order, extra_info)
        -> (xt, zs[, hspaces[, skipconns]])                                      reference :147-323

Two execution paths, both on the libaedit kernels:
  * fused path (default; no h-space / skip taps requested): one batched U-Net launch per step for the
    uncond+cond CFG rows, CFG combine + scheduler update fused in one kernel (ae_cfg_inv_step / ae_cfg_rev_step),
    no per-step host synchronisation.  The forward process can additionally batch `forward_batch` timesteps per
    launch (every U-Net input of the forward process is sampled directly from x0); with
    forward_batch=1 the loop is step-sequential and reproduces the reference's data flow exactly, including the
    bit-exact replay invariant.
  * general path (taps requested): the reference's per-step structure through the wrapper methods.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple, Union

import torch
from tqdm import tqdm

from ..models import PipelineWrapper

DEFAULT_FORWARD_BATCH = int(os.environ.get("AEDIT_FORWARD_BATCH", "50"))
USE_CUDA_GRAPHS = os.environ.get("AEDIT_CUDA_GRAPH", "1") != "0"
# Forward / reverse overlap (see _PendingForward): 1 = on (default), 0 = both processes on the caller's stream.
OVERLAP = os.environ.get("AEDIT_OVERLAP", "1") != "0"
# Which graph variant the reverse lane replays while forward chunks are still running: "adaptive" (default) = the
# shared-SM variant (unet.GraphedForward lane 2) until the forward lane's last event has fired, then the solo variant;
# "solo" / "shared" pin one variant (A/B measurements).  All variants give identical bits.
REV_VARIANT = os.environ.get("AEDIT_REV_VARIANT", "adaptive")
REV_LOOKAHEAD = int(os.environ.get("AEDIT_REV_LOOKAHEAD", "2"))


class _PendingForward:
    """Book-keeping of a forward process whose timestep chunks were enqueued on the model's forward lane.

    The forward process is throughput-bound (B = 2 x forward_batch rows per launch fill the machine) and its
    timesteps are independent; the reverse process is a strictly sequential chain of sub-wave launches
    that leaves most SMs idle.  So the two run CONCURRENTLY: the forward chunks are enqueued on a side stream in the
    order the reverse process will consume them, each followed by an event; `inversion_reverse_process` — called right
    after, exactly like reference main_run.py:128-158 — runs on a second, high-priority side stream and waits only
    for the events of the chunks whose rows (`zs[idx]`, `xts[tstart]`) it is about to read.  The caller's stream waits
    for the whole forward lane when `inversion_forward_process` returns and for the reverse lane when
    `inversion_reverse_process` returns, so every other consumer of the returned tensors sees ordinary stream
    semantics.  The fast path is taken only if the tensors handed to the reverse process are the ones the forward
    process returned, unmodified (same storage, same torch version counters) — anything else falls back to plain
    stream order.  Results are bit-identical to the non-overlapped execution (same kernels, same batches; tested)."""
    __slots__ = ("zs_ptr", "zs_ver", "xts_ptr", "xts_ver", "chunks", "eta_key", "N", "setup_event")

    def __init__(self):
        self.chunks = []          # (idx_lo, idx_hi, event): rows zs[lo..hi], xts[lo..hi] are final once event fired

    def matches(self, zs, xT, eta_key) -> bool:
        try:
            return (zs.data_ptr() == self.zs_ptr and zs._version == self.zs_ver and xT.data_ptr() == self.xts_ptr
                    and xT._version == self.xts_ver and zs.shape[0] <= self.N and xT.shape[0] == self.N + 1
                    and eta_key == self.eta_key)
        except RuntimeError:      # inference tensors carry no version counter: cannot prove they are unmodified
            return False

    def event_for(self, idx):
        for lo, hi, ev in self.chunks:
            if lo <= idx <= hi:
                return ev
        return None


def _lane(model, name: str):
    """Side streams of a model: 'fwd' (default priority) and 'rev' (highest priority)."""
    lanes = model.__dict__.setdefault("_lanes", {})
    st = lanes.get(name)
    if st is None:
        st = torch.cuda.Stream(device=model.device, priority=-1 if name == "rev" else 0)
        lanes[name] = st
    return st


DEFAULT_HEAD_CHUNK = int(os.environ.get("AEDIT_HEAD_CHUNK", "10"))


def _chunk_plan(N: int, tb: int, hint: Optional[int], head: Optional[int] = None) -> List[Tuple[int, int]]:
    """(pos0, count) of the forward-process chunks in launch order.  Without a hint: loop order, boundaries at
    multiples of `tb`.  With hint = tstart of the reverse process that follows: the rows idx (= N - pos - 1) are cut
    so that row `hint` is the TOP of a chunk — the reverse process starts from xts[hint] and then consumes zs[hint-1],
    zs[hint-2], ..., so one chunk is all it has to wait for — and the chunks are launched downwards from there, the
    rows above `hint` last.  The first chunk holds only `head` rows (default AEDIT_HEAD_CHUNK = 10): it is all the
    reverse process waits for before its first step, and the rows it consumes next are delivered by the following
    chunk while it works through those.  A remainder shorter than tb/4 is merged into its neighbour.  The plan depends
    only on (N, tb, hint, head), never on whether the lanes overlap, so both modes run identical batches
    (bit-identical results)."""
    if tb <= 1 or hint is None or not (0 <= hint <= N):
        return [(p, min(tb, N - p)) for p in range(0, N, tb)]
    head = DEFAULT_HEAD_CHUNK if head is None else head
    top = min(hint, N - 1)
    down, up = [], []                       # (lo, hi) in idx space
    hi = top
    if 0 < head < tb and top - head + 1 > 0:
        down.append((top - head + 1, top))
        hi = top - head
    while hi >= 0:
        lo = max(0, hi - tb + 1)
        if lo > 0 and lo < max(1, tb // 4):
            lo = 0
        down.append((lo, hi))
        hi = lo - 1
    lo = top + 1
    while lo <= N - 1:
        hi = min(N - 1, lo + tb - 1)
        if N - 1 - hi < max(1, tb // 4):
            hi = N - 1
        up.append((lo, hi))
        lo = hi + 1
    return [(N - 1 - hi, hi - lo + 1) for lo, hi in down + up]


def _unet_eval(model, x_in, t_in, text, slot, cl, slot_key=None, lane=0):
    """One batched U-Net evaluation, through a cached CUDA graph unless AEDIT_CUDA_GRAPH=0.  lane 1 = the reverse
    lane's graph: own workspaces, kernel nodes captured with the highest launch priority (unet.GraphedForward)."""
    eng = model.engine
    slot_arg = slot if text is not None else None
    if USE_CUDA_GRAPHS and x_in.is_cuda:
        g = eng.graphed(x_in.shape[0], x_in.shape[2], x_in.shape[3], text, slot_arg, cl, slot_key=slot_key, lane=lane)
        model.graph_replays = getattr(model, "graph_replays", 0) + 1
        model.graph_kernels = getattr(model, "graph_kernels", 0) + g.kernels
        return g(x_in, t_in, cl)
    return eng.forward(x_in, t_in, text=text, slot_map=slot_arg, class_labels=cl)


def _gaussian_blur_k15_s1(x: torch.Tensor) -> torch.Tensor:
    """torchvision.transforms.functional.gaussian_blur(x, kernel_size=15, sigma=1) (reflect padding, separable
    kernel) — used once per run on the multi-prompt cfg / mask maps (inversion_utils.py:49,197-198)."""
    from torchvision.transforms import functional as TF
    return TF.gaussian_blur(x, kernel_size=15, sigma=1)


def _build_cfg_maps(batch_size, shape, cfg_scales, cutoff_points, device, dtype, prompts=None, masks_too=False):
    """inversion_utils.py:29-51 (forward) / :177-200 (reverse).  Mutates cfg_scales in place like the reference
    (`cfg_scales *= batch_size`)."""
    cfg_scales_tensor = torch.ones((batch_size, *shape), device=device, dtype=dtype)
    masks = torch.ones((batch_size, *shape), device=device, dtype=dtype) if masks_too else None
    if batch_size > 1:
        if cutoff_points is None:
            cutoff_points = [i * 1 / batch_size for i in range(1, batch_size)]
        if len(cfg_scales) == 1:
            cfg_scales *= batch_size
        elif len(cfg_scales) < batch_size:
            raise ValueError("Not enough target CFG scales")
        cutoff_points = [int(x * cfg_scales_tensor.shape[2]) for x in cutoff_points]
        cutoff_points = [0, *cutoff_points, cfg_scales_tensor.shape[2]]
        for i, (start, end) in enumerate(zip(cutoff_points[:-1], cutoff_points[1:])):
            cfg_scales_tensor[i, :, end:] = 0
            cfg_scales_tensor[i, :, :start] = 0
            if masks_too:
                masks[i, :, end:] = 0
                masks[i, :, :start] = 0
            cfg_scales_tensor[i] *= cfg_scales[i]
            if (not masks_too) and prompts is not None and prompts[i] == "":
                cfg_scales_tensor[i] = 0
        cfg_scales_tensor = _gaussian_blur_k15_s1(cfg_scales_tensor)
        if masks_too:
            masks = _gaussian_blur_k15_s1(masks)
    else:
        cfg_scales_tensor *= cfg_scales[0]
    return cfg_scales_tensor.contiguous(), (masks.contiguous() if masks_too else None)


def _cat_text(model: PipelineWrapper, uncond, cond):
    """Stack the uncond row(s) and the P cond rows of (hidden_states, class_labels, mask) into one text batch,
    right-padding the token axis (padding keys get a large negative bias, i.e. exactly zero softmax weight)."""
    streams_u, masks_u, cl_u = model._text_for(*uncond)
    if cond is None:
        return streams_u, masks_u, cl_u
    streams_c, masks_c, cl_c = model._text_for(*cond)
    streams, masks = [], []
    for su, sc, mu, mc in zip(streams_u, streams_c, masks_u, masks_c):
        L = max(su.shape[1], sc.shape[1])
        need_mask = (mu is not None) or (mc is not None) or su.shape[1] != sc.shape[1]

        def pad(s, m):
            n, l = s.shape[0], s.shape[1]
            if m is None:
                m = torch.ones(n, l, device=s.device)
            m = m.to(torch.float32)
            if l < L:
                s = torch.cat([s, s.new_zeros(n, L - l, s.shape[2])], 1)
                # padded slots: mask value chosen so that (1-m)*-10000 becomes a huge negative bias
                m = torch.cat([m, m.new_full((n, L - l), -1.0e26)], 1)
            return s, m
        su2, mu2 = pad(su, mu)
        sc2, mc2 = pad(sc, mc)
        streams.append(torch.cat([su2, sc2], 0))
        masks.append(torch.cat([mu2, mc2], 0) if need_mask else None)
    cl = None if cl_u is None else torch.cat([cl_u, cl_c], 0)
    return streams, masks, cl


def _loop_text(model: PipelineWrapper, neg_prompts, prompts):
    """Text conditioning of one loop call: (TextCache or None, class-label rows or None) for the rows
    [uncond, cond_1..cond_P].  Cached per prompt strings — the text encoders are deterministic, so re-encoding the
    same prompts for every call (as the reference does) would only rebuild identical K/V tensors (and force the
    CUDA graphs bound to them to be re-captured)."""
    enc = model.encode_text
    key = (tuple(neg_prompts), None if prompts is None else tuple(prompts), id(getattr(enc, "__self__", enc)))
    cache = model.__dict__.setdefault("_loop_text_cache", {})
    hit = cache.get(key)
    if hit is None:
        if len(cache) > 8:
            cache.clear()
            model.engine.evict_graphs(model.live_texts() if hasattr(model, "live_texts") else ())
        uncond = model.encode_text(list(neg_prompts), negative=True)
        cond = None if prompts is None else model.encode_text(list(prompts))
        streams, masks, cl = _cat_text(model, uncond, cond)
        text = model.engine.prepare_text(streams, masks) if streams else None
        hit = (text, cl, (streams, masks))
        cache[key] = hit
    return hit[0], hit[1]


def _loop_text_cached(model: PipelineWrapper, neg_prompts, prompts) -> bool:
    """True if _loop_text would be a cache hit (no text-encoder / K|V-projection launches)."""
    enc = model.encode_text
    key = (tuple(neg_prompts), None if prompts is None else tuple(prompts), id(getattr(enc, "__self__", enc)))
    return key in model.__dict__.get("_loop_text_cache", {})


def _t_to_idx(timesteps):
    if timesteps[0].dtype == torch.int64:
        return {int(v): k for k, v in enumerate(timesteps)}
    return {float(v): k for k, v in enumerate(timesteps)}


def inversion_forward_process(model: PipelineWrapper,
                              x0: torch.Tensor,
                              etas: Optional[float] = None,
                              prog_bar: bool = False,
                              prompts: List[str] = [""],
                              cfg_scales: List[float] = [3.5],
                              num_inference_steps: int = 50,
                              cutoff_points: Optional[List[float]] = None,
                              numerical_fix: bool = False,
                              extract_h_space: bool = False,
                              extract_skipconns: bool = False,
                              duration: Optional[float] = None,
                              first_order: bool = False,
                              forward_batch: Optional[int] = None,
                              noise: Optional[torch.Tensor] = None,
                              group=None,
                              reverse_hint: Optional[int] = None) -> Tuple:
    """Extensions over the reference signature (all default to the reference behaviour): `forward_batch` timesteps
    per U-Net launch, explicit `noise`, `reverse_hint` — the `tstart` the following reverse process will
    use (default N // 2, the ratio of main_run.py's own defaults 200 / 100); it only orders the chunk launches so the
    reverse process can start early (see _PendingForward), never changes a result — and `group` — a torch.distributed process group over which the
    timestep chunks of ONE clip are sharded: chunk k runs on group rank k % world, `xts` is
    broadcast from rank 0 once, and `zs` / `xts` are merged at the end by one all-gather of each rank's owned rows
    (parallel.merge_owned_rows_: every row is owned by exactly one rank, every rank returns the full tensors)."""
    if len(prompts) > 1 and extract_h_space:
        raise NotImplementedError("How do you split cfg_scales for hspace? TODO")
    if extract_h_space or extract_skipconns:
        return _forward_general(model, x0, etas, prog_bar, prompts, cfg_scales, num_inference_steps, cutoff_points,
                                numerical_fix, extract_h_space, extract_skipconns, duration, first_order)

    have_cond = len(prompts) > 1 or prompts[0] != ""
    P = len(prompts) if have_cond else 0
    cfg_map = None
    if have_cond:
        cfg_map, _ = _build_cfg_maps(P, x0.shape[1:], cfg_scales, cutoff_points, model.device, x0.dtype, prompts)
    text, cl = _loop_text(model, [""], prompts if have_cond else None)
    sched = model.model.scheduler
    timesteps = sched.timesteps.to(model.device)
    N = num_inference_steps
    if type(etas) in [int, float]:
        etas = [etas] * sched.num_inference_steps
    with torch.inference_mode(False):      # normal tensors: their version counters guard the overlap fast path
        xts = model.sample_xts_from_x0(x0, num_inference_steps=N, noise=noise)
        zs = torch.zeros(size=model.get_noise_shape(x0, N), device=model.device)
    extra_info = [None] * len(zs)
    model.setup_extra_inputs(x0, init_timestep=timesteps[0], audio_end_in_s=duration)
    model.__dict__.pop("_pending_forward", None)

    tb = forward_batch if forward_batch is not None else DEFAULT_FORWARD_BATCH
    tb = max(1, min(int(tb), N))
    g_rank, g_ws = (0, 1)
    if group is not None:
        from .. import parallel as _par
        g_rank, g_ws = _par.world(group)
        if g_ws > 1:
            if tb == 1:
                raise ValueError("timestep sharding needs forward_batch > 1 (the step-sequential mode chains x_t)")
            _par.broadcast_(xts, 0, group)            # one noise draw for all ranks
    owned: List[int] = []
    n_el = x0[0].numel()
    xt_src = xts.clone() if tb > 1 else xts      # batched: every U-Net input is the directly sampled x_t (F8)
    ts_cpu = sched.timesteps_cpu
    model.sched_table.set_etas(etas)
    overlap = OVERLAP and USE_CUDA_GRAPHS and x0.is_cuda and tb > 1 and g_ws == 1 and not prog_bar
    # the same plan with and without a process group: a chunk's bits depend on its batch, so sharded == single-GPU
    plan = _chunk_plan(N, tb, N // 2 if reverse_hint is None else int(reverse_hint))
    it = tqdm(plan) if prog_bar else plan
    # everything a chunk needs from the host is staged BEFORE the first launch: a pageable host->device copy on a busy
    # stream blocks the host until the stream drains, which would serialise the launches of the two lanes
    slots = {}
    for count in {c for _, c in plan}:
        if P > 0:
            sl = torch.cat([torch.zeros(count, dtype=torch.int32), (1 + torch.arange(P, dtype=torch.int32)).repeat(count)])
        else:
            sl = torch.zeros(count, dtype=torch.int32)
        sl = sl.to(model.device)
        slots[count] = (sl, None if cl is None else cl.index_select(0, sl.long()))
    cur = torch.cuda.current_stream() if x0.is_cuda else None
    pend = None
    if overlap:
        pend = _PendingForward()
        lane = _lane(model, "fwd")
        # xts (ae_sample_xts) and the zs fill were enqueued on the caller's stream: the forward lane waits for them here,
        # and the reverse lane — which on the fast path never joins the caller's stream — waits for this event before it
        # reads xts[tstart] (row N is written by ae_sample_xts only, no forward chunk finalises it)
        pend.setup_event = torch.cuda.Event()
        pend.setup_event.record(cur)
        lane.wait_stream(cur)
    for chunk_no, (pos0, count) in enumerate(it):
        if g_ws > 1:
            if chunk_no % g_ws != g_rank:
                continue
            owned.extend(range(N - pos0 - count, N - pos0))
        with torch.cuda.stream(lane) if overlap else _NullCtx():
            # loop position pos <-> idx = N - pos - 1 (inversion_utils.py:75); U-Net input xts[idx+1] = xts[N - pos]
            src_rows = torch.arange(N - pos0, N - pos0 - count, -1, device=model.device)
            xt_b = xt_src.index_select(0, src_rows)                                    # [count, C, H, W]
            t_b = timesteps[pos0:pos0 + count]
            slot, cl_b = slots[count]
            if P > 0:
                x_in = torch.cat([xt_b, xt_b.repeat_interleave(P, 0)], 0)
                t_in = torch.cat([t_b, t_b.repeat_interleave(P)], 0)
            else:
                x_in, t_in = xt_b, t_b
            eps = _unet_eval(model, x_in, t_in, text, slot, cl_b, slot_key=("fwd", count, P))
            eta = float(etas[N - pos0 - 1])
            model.k_cfg_inv_step(pos0, count, eta, eps, eps[count:] if P > 0 else None, P, cfg_map, xt_src, xts, zs,
                                 numerical_fix)
            if overlap:
                lo, hi = N - pos0 - count, N - pos0 - 1
                if lo == 0:
                    zs[0] = torch.zeros_like(zs[0])     # inversion_utils.py:133, in lane order
                ev = torch.cuda.Event()
                ev.record(lane)
                pend.chunks.append((lo, hi, ev))
    if g_ws > 1:
        _par.merge_owned_rows_(zs, owned, group)
        _par.merge_owned_rows_(xts[:N], owned, group)
    xt = xts[1][None] if N >= 1 else x0                 # the reference returns the last loop's xt = xts[1]
    if overlap:
        cur.wait_stream(lane)                           # ordinary stream semantics for every other consumer
        pend.zs_ptr, pend.zs_ver, pend.xts_ptr, pend.xts_ver = zs.data_ptr(), zs._version, xts.data_ptr(), xts._version
        pend.eta_key, pend.N = model.sched_table._eta_key, N
        model.__dict__["_pending_forward"] = pend
    else:
        zs[0] = torch.zeros_like(zs[0])                 # inversion_utils.py:133
    return xt, zs, xts, extra_info


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def inversion_reverse_process(model: PipelineWrapper,
                              xT: torch.Tensor,
                              tstart: torch.Tensor,
                              fix_alpha: float = 0.1,
                              etas: float = 0,
                              prompts: List[str] = [""],
                              neg_prompts: List[str] = [""],
                              cfg_scales: Optional[List[float]] = None,
                              prog_bar: bool = False,
                              zs: Optional[List[torch.Tensor]] = None,
                              cutoff_points: Optional[List[float]] = None,
                              hspace_add: Optional[torch.Tensor] = None,
                              hspace_replace: Optional[torch.Tensor] = None,
                              skipconns_replace: Optional[Dict[int, torch.Tensor]] = None,
                              zero_out_resconns: Optional[Union[int, List]] = None,
                              extract_h_space: bool = False,
                              extract_skipconns: bool = False,
                              duration: Optional[float] = None,
                              first_order: bool = False,
                              extra_info: Optional[List] = None,
                              trace: Optional[List] = None) -> Tuple:
    """`trace` (optional list) receives a clone of x_t after every step — a debugging / test hook that the
    reference does not have."""
    taps = (hspace_add is not None or hspace_replace is not None or skipconns_replace is not None
            or zero_out_resconns is not None or extract_h_space or extract_skipconns)
    if taps or not prompts:
        return _reverse_general(model, xT, tstart, fix_alpha, etas, prompts, neg_prompts, cfg_scales, prog_bar, zs,
                                cutoff_points, hspace_add, hspace_replace, skipconns_replace, zero_out_resconns,
                                extract_h_space, extract_skipconns, duration, first_order, extra_info)
    P = len(prompts)
    sched = model.model.scheduler
    N = sched.num_inference_steps
    if etas is None:
        etas = 0
    if type(etas) in [int, float]:
        etas = [etas] * N
    assert len(etas) == N
    n = zs.shape[0]
    tmax = int(tstart.max())
    eta_key = tuple(float(e) for e in etas)

    # ---- lanes (see _PendingForward): run on the high-priority reverse lane; wait per chunk when the forward process
    # that produced (zs, xT) is still in flight, else behind the caller's stream
    lanes = OVERLAP and USE_CUDA_GRAPHS and xT.is_cuda
    pend = model.__dict__.pop("_pending_forward", None) if lanes else None
    cur = torch.cuda.current_stream() if xT.is_cuda else None
    if pend is not None and not (pend.matches(zs, xT, eta_key) and _loop_text_cached(model, neg_prompts, prompts)
                                 and bool((tstart == tmax).all())):
        pend = None
    if lanes:
        lane = _lane(model, "rev")
        if pend is None:
            lane.wait_stream(cur)
        else:
            model.overlap_hits = getattr(model, "overlap_hits", 0) + 1
            lane.wait_event(pend.setup_event)
    waited = set()

    def need(idx):
        """Make the reverse lane wait for the forward chunk that finalises row idx of zs / xts."""
        if pend is None or idx >= pend.N:       # row N of xts is never rewritten by the forward process
            return
        ev = pend.event_for(idx)
        if ev is not None and id(ev) not in waited:
            waited.add(id(ev))
            lane.wait_event(ev)

    with torch.cuda.stream(lane) if lanes else _NullCtx():
        text, cl = _loop_text(model, neg_prompts, prompts)
        cfg_map, masks = _build_cfg_maps(P, xT.shape[1:], cfg_scales, cutoff_points, model.device, xT.dtype,
                                         masks_too=True)
        need(tmax)
        xt = xT[tmax].unsqueeze(0).to(torch.float32).contiguous().clone()
        ts_cpu = sched.timesteps_cpu[-n:]
        model.setup_extra_inputs(xt, extra_info=extra_info, init_timestep=ts_cpu[0], audio_end_in_s=duration)
        rows = 1 + P
        model.sched_table.set_etas(etas)
        slot = torch.arange(rows, dtype=torch.int32, device=model.device)
        zs = zs.contiguous()
        it = range(n)
        if prog_bar:
            it = tqdm(it)
        x_in = torch.empty((rows, *xt.shape[1:]), device=model.device, dtype=torch.float32)
        # While forward chunks are still running, the steps replay the shared-SM graph variant and are enqueued at
        # most REV_LOOKAHEAD steps ahead of the device, so that the switch to the solo variant happens (up to that
        # look-ahead) when the forward lane has actually drained — the host polls its last event.
        # "forward lane busy" = ANY forward-process work still queued on the model's forward lane: this clip's remaining
        # chunks, or (edit_clips_pipelined) the next clip's forward process that was enqueued ahead of this call
        fwd_lane = model.__dict__.get("_lanes", {}).get("fwd") if pend is not None else None
        contended = fwd_lane is not None and REV_VARIANT != "solo"
        in_flight = []
        for k in it:
            t = int(ts_cpu[k])
            pos = N - n + k
            idx = n - k - 1                                                      # inversion_utils.py:222-224
            if contended and REV_VARIANT == "adaptive":
                if len(in_flight) >= REV_LOOKAHEAD:
                    in_flight.pop(0).synchronize()
                if fwd_lane.query():
                    contended = False
            variant = 0 if not lanes else (2 if (contended or REV_VARIANT == "shared") else 1)
            if variant == 1 and REV_VARIANT == "adaptive":
                variant = model.engine.solo_lane(rows, x_in.shape[2], x_in.shape[3], text, ("rev", P), cl is not None)
            x_in.copy_(xt.expand(rows, -1, -1, -1))
            t_in = torch.full((rows,), t, dtype=torch.int64, device=model.device)
            eps = _unet_eval(model, x_in, t_in, text, slot, cl, slot_key=("rev", P), lane=variant)
            apply_fix = ((tstart.max() - tstart) > k)
            fa = None
            xT_fix = None
            if apply_fix.any():                                                  # inversion_utils.py:308-315
                fa = [float(v) for v in (apply_fix * fix_alpha).to(torch.float32)]
                xT_fix = xT[tmax - k - 1].to(torch.float32).contiguous()
            out = torch.empty_like(xt)
            need(idx)
            model.k_cfg_rev_step(pos, float(etas[idx]), eps, eps[1:], P, cfg_map, xt, zs[idx], out, masks=masks,
                                 fix_alpha=fa, xT_fix=xT_fix)
            xt = out
            if contended and REV_VARIANT == "adaptive":
                ev = torch.cuda.Event()
                ev.record(lane)
                in_flight.append(ev)
            if trace is not None:
                trace.append(xt.clone())
    if lanes:
        cur.wait_stream(lane)
        xt.record_stream(cur)                   # allocated on the lane, handed to the caller's stream
        if trace is not None:
            for t_ in trace:
                t_.record_stream(cur)
    return xt, zs


# ------------------------------------------------------------------------------------------------------------------
# Clip queue with cross-clip pipelining.  Within one clip the reverse process can only overlap the TAIL of its own
# forward process (it needs zs[tstart-1..] first) and then runs alone for most of its steps, a latency-bound chain that
# leaves most SMs idle — exactly the window in which the NEXT clip's throughput-bound forward process can run.  So the
# queue enqueues forward(i+1) on the forward lane BEFORE it drives reverse(i) on the reverse lane; every call is the
# unchanged single-clip drop-in function (same kernels, same batches, same bits as calling them one clip at a time).
# ------------------------------------------------------------------------------------------------------------------
def edit_clips_pipelined(model: PipelineWrapper, x0s, src_prompts: List[str], tgt_prompts: List[str], tstart: int,
                         cfg_src: float = 3.0, cfg_tar: float = 12.0, num_inference_steps: int = 200, etas: float = 1.0,
                         neg_prompts: List[str] = [""], numerical_fix: bool = True, forward_batch: Optional[int] = None,
                         noises=None, on_result=None, group: int = 1) -> List[torch.Tensor]:
    """Edit a queue of clips (main_run.py:127-160 per clip: inversion_forward_process, then inversion_reverse_process
    from `tstart`) with the forward process of the NEXT clip(s) running concurrently with the reverse process of the
    current one(s).  x0s: iterable of [1,C,H,W] latents (device tensors, or pinned host tensors that are copied in when
    their turn comes); noises: optional per-clip [N,C,H,W]; on_result(i, w): optional callback as soon as clip i's edit
    is enqueued (e.g. an asynchronous device-to-host copy).  Returns the edited latents [1,C,H,W] in queue order.

    group = 1: every call is the unchanged single-clip drop-in function (same kernels, batches and bits as editing the
    clips one at a time); forward(i+1) is enqueued on the forward lane BEFORE reverse(i) is driven on the reverse lane.
    group = K > 1: K clips per reverse launch (inversion_*_batched, B = K*(1+P) rows): the latency-bound reverse chain
    then carries K clips per step, and the next group's forward process fills the SMs it leaves idle."""
    clips = list(x0s)
    n = len(clips)
    N = num_inference_steps
    # text conditioning of both prompt sets is prepared up front on the caller's stream: its GEMMs share workspaces with
    # the forward lane's kernels and must not run beside them
    _loop_text(model, [""], list(src_prompts) if (len(src_prompts) > 1 or src_prompts[0] != "") else None)
    _loop_text(model, list(neg_prompts), list(tgt_prompts))
    if group <= 1:
        ts = torch.tensor([int(tstart)], dtype=torch.int)

        def forward(i):
            x0 = clips[i].to(model.device, non_blocking=True)
            _, zs, xts, _ = inversion_forward_process(model, x0, etas=etas, prompts=list(src_prompts), cfg_scales=[cfg_src],
                                                      num_inference_steps=N, numerical_fix=numerical_fix,
                                                      forward_batch=forward_batch, reverse_hint=int(tstart),
                                                      noise=None if noises is None else noises[i])
            return zs, xts, model.__dict__.pop("_pending_forward", None)

        out: List[torch.Tensor] = []
        nxt = forward(0) if n else None
        for i in range(n):
            zs, xts, pend = nxt
            nxt = forward(i + 1) if i + 1 < n else None          # enqueued on the forward lane ahead of reverse(i)
            if pend is not None:
                model.__dict__["_pending_forward"] = pend
            w, _ = inversion_reverse_process(model, xT=xts, tstart=ts, etas=etas, prompts=list(tgt_prompts),
                                             neg_prompts=list(neg_prompts), cfg_scales=[cfg_tar], zs=zs[:int(tstart)])
            out.append(w)
            if on_result is not None:
                on_result(i, w)
        return out

    # ---- groups of K clips: forward lane / reverse lane, one event per group
    cur = torch.cuda.current_stream()
    f_lane, r_lane = _lane(model, "fwd"), _lane(model, "rev")
    groups = [list(range(g, min(n, g + group))) for g in range(0, n, group)]

    def forward_group(ids):
        f_lane.wait_stream(cur)
        with torch.cuda.stream(f_lane):
            x = torch.cat([clips[i].to(model.device, non_blocking=True) for i in ids], 0)
            nz = None if noises is None else torch.stack([noises[i] for i in ids])
            _, zs, xts = inversion_forward_process_batched(model, x, etas=etas, prompts=list(src_prompts),
                                                           cfg_scales=[cfg_src], num_inference_steps=N,
                                                           numerical_fix=numerical_fix, forward_batch=forward_batch, noise=nz)
            ev = torch.cuda.Event()
            ev.record(f_lane)
        for t_ in (zs, xts):
            t_.record_stream(r_lane)              # allocated on the forward lane, consumed on the reverse lane
        return zs, xts, ev

    out = [None] * n
    nxt = forward_group(groups[0]) if groups else None
    for gi, ids in enumerate(groups):
        zs, xts, ev = nxt
        more = gi + 1 < len(groups)
        nxt = forward_group(groups[gi + 1]) if more else None       # ahead of this group's reverse process
        r_lane.wait_event(ev)
        with torch.cuda.stream(r_lane):
            w, _ = inversion_reverse_process_batched(model, xts, int(tstart), etas=etas, prompts=list(tgt_prompts),
                                                     neg_prompts=list(neg_prompts), cfg_scales=[cfg_tar],
                                                     zs=zs[:, :int(tstart)], graph_lane=2 if more else 1)
            for k, i in enumerate(ids):
                out[i] = w[k:k + 1]
                if on_result is not None:
                    on_result(i, out[i])          # runs in reverse-lane stream order (e.g. an async D2H copy)
        w.record_stream(cur)
    cur.wait_stream(r_lane)
    cur.wait_stream(f_lane)
    return out


# ------------------------------------------------------------------------------------------------------------------
# Multi-clip batches (BASELINE configs[2]: "batch of 32 clips sharded across 8 GPUs").  The reference has
# no clip axis — its batch dimension means "prompts of one clip" (inversion_utils.py:78,88,252; get_noise_shape,
# models.py:60-65) — so these are NEW entry points; the single-clip drop-in signatures above are untouched.  Every U-Net
# launch carries B = K * (1 + P) rows (K clips x [uncond, P prompts]): the reverse process of one clip is a sequential
# chain of sub-wave launches that leaves most SMs idle, K clips per launch turn that idle width into throughput.
# Per clip the arithmetic is the same kernels on the same data as the single-clip functions (a row's bits depend on the
# launch's batch size only through the fixed, batch-size-dependent split-K plan of the small-M GEMMs).
# ------------------------------------------------------------------------------------------------------------------
def _clip_prompts(prompts, K: int) -> List[List[str]]:
    """prompts: List[str] shared by all clips, or List[List[str]] with one list per clip (equal lengths)."""
    if len(prompts) and isinstance(prompts[0], (list, tuple)):
        if len(prompts) != K:
            raise ValueError(f"{len(prompts)} prompt lists for {K} clips")
        if len({len(p) for p in prompts}) != 1:
            raise ValueError("every clip needs the same number of prompts")
        return [list(p) for p in prompts]
    return [list(prompts) for _ in range(K)]


def _dev_i32(model: PipelineWrapper, values) -> torch.Tensor:
    """Small int32 device constant, uploaded ONCE per model and value list: a pageable host->device copy blocks the host
    until the stream it is enqueued on has drained, which would serialise the lanes of the clip queue."""
    cache = model.__dict__.setdefault("_i32_consts", {})
    key = tuple(values)
    t = cache.get(key)
    if t is None:
        if len(cache) > 64:
            cache.clear()
        t = torch.tensor(list(values), dtype=torch.int32, device=model.device)
        cache[key] = t
    return t


def _batched_text(model: PipelineWrapper, neg_prompts: List[str], per_clip: List[List[str]]):
    """Text rows of a multi-clip launch: row 0 = the negative / unconditional prompt, then the DISTINCT conditional prompts
    (clips sharing a prompt share its K/V rows).  Returns (TextCache, class-label rows, {prompt: row})."""
    uniq: List[str] = []
    for ps in per_clip:
        for q in ps:
            if q not in uniq:
                uniq.append(q)
    text, cl = _loop_text(model, neg_prompts, uniq if uniq else None)
    return text, cl, {q: 1 + i for i, q in enumerate(uniq)}


def inversion_forward_process_batched(model: PipelineWrapper, x0s: torch.Tensor, etas: float = 1.0,
                                      prompts=("",), cfg_scales: List[float] = [3.5], num_inference_steps: int = 50,
                                      cutoff_points: Optional[List[float]] = None, numerical_fix: bool = True,
                                      forward_batch: Optional[int] = None, noise: Optional[torch.Tensor] = None
                                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """inversion_forward_process (reference :8-144) for K clips at once.  x0s: [K, C, H, W]; noise (optional):
    [K, N, C, H, W]; prompts: shared List[str] or one list per clip.  Returns (xt [K,1,C,H,W], zs [K,N,C,H,W],
    xts [K,N+1,C,H,W]) — per clip what the single-clip function returns."""
    K = x0s.shape[0]
    N = num_inference_steps
    per_clip = _clip_prompts(list(prompts), K)
    have_cond = len(per_clip[0]) > 1 or per_clip[0][0] != ""
    P = len(per_clip[0]) if have_cond else 0
    cfg_map = None
    if have_cond:
        cfg_map, _ = _build_cfg_maps(P, x0s.shape[1:], list(cfg_scales), cutoff_points, model.device, x0s.dtype, per_clip[0])
    text, cl, row_of = _batched_text(model, [""], per_clip if have_cond else [[]] * K)
    sched = model.model.scheduler
    timesteps = sched.timesteps.to(model.device)
    if type(etas) in [int, float]:
        etas = [etas] * sched.num_inference_steps
    xts = torch.stack([model.sample_xts_from_x0(x0s[k:k + 1], num_inference_steps=N,
                                                noise=None if noise is None else noise[k]) for k in range(K)])
    zs = torch.zeros((K, *model.get_noise_shape(x0s, N)), device=model.device)
    model.sched_table.set_etas(etas)
    tb = forward_batch if forward_batch is not None else DEFAULT_FORWARD_BATCH
    tb = max(1, min(int(tb) // K if tb > 1 else 1, N))           # keep rows per launch ~ 2 * forward_batch
    xt_src = xts.clone() if tb > 1 else xts
    rows_per_clip = 1 + P
    for pos0 in range(0, N, tb):
        count = min(tb, N - pos0)
        src_rows = torch.arange(N - pos0, N - pos0 - count, -1, device=model.device)
        t_b = timesteps[pos0:pos0 + count]
        xs, ts_, sl = [], [], []
        for k in range(K):
            xt_b = xt_src[k].index_select(0, src_rows)
            xs += [xt_b] + ([xt_b.repeat_interleave(P, 0)] if P else [])
            ts_ += [t_b] + ([t_b.repeat_interleave(P)] if P else [])
            sl += [0] * count + [row_of[q] for _ in range(count) for q in per_clip[k]] if P else [0] * count
        x_in, t_in = torch.cat(xs, 0), torch.cat(ts_, 0)
        slot = _dev_i32(model, sl)
        cl_b = None if cl is None else cl.index_select(0, slot.long())
        eps = _unet_eval(model, x_in, t_in, text, slot, cl_b, slot_key=("fwdK", K, count, tuple(sl)))
        eta = float(etas[N - pos0 - 1])
        blk = count * rows_per_clip
        for k in range(K):
            e = eps[k * blk:(k + 1) * blk]
            model.k_cfg_inv_step(pos0, count, eta, e, e[count:] if P else None, P, cfg_map, xt_src[k], xts[k], zs[k],
                                 numerical_fix)
    zs[:, 0] = 0                                                      # inversion_utils.py:133
    return xts[:, 1:2], zs, xts


def inversion_reverse_process_batched(model: PipelineWrapper, xT: torch.Tensor, tstart: int, etas: float = 1.0,
                                      prompts=("",), neg_prompts: List[str] = [""],
                                      cfg_scales: Optional[List[float]] = None, zs: Optional[torch.Tensor] = None,
                                      cutoff_points: Optional[List[float]] = None, graph_lane: Optional[int] = None
                                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """inversion_reverse_process (reference :147-323) for K clips at once, one tstart for all prompts and clips.
    xT: [K, N+1, C, H, W] (the xts of the forward process), zs: [K, n, C, H, W] (= zs[:, :tstart]).
    graph_lane: which captured graph variant to replay (unet.GraphedForward: 1 = the lane has the machine to itself,
    2 = rings / PDL sized for sharing the SMs with a forward-process grid on another stream); same bits either way.
    Returns (edited latents [K, C, H, W], zs)."""
    K = xT.shape[0]
    per_clip = _clip_prompts(list(prompts), K)
    P = len(per_clip[0])
    sched = model.model.scheduler
    N = sched.num_inference_steps
    if type(etas) in [int, float]:
        etas = [etas] * N
    n = zs.shape[1]
    tmax = int(tstart)
    text, cl, row_of = _batched_text(model, neg_prompts, per_clip)
    cfg_map, _ = _build_cfg_maps(P, xT.shape[2:], list(cfg_scales), cutoff_points, model.device, xT.dtype, masks_too=True)
    model.sched_table.set_etas(etas)
    rows = 1 + P
    sl = []
    for k in range(K):
        sl += [0] + [row_of[q] for q in per_clip[k]]
    slot = _dev_i32(model, sl)
    cl_b = None if cl is None else cl.index_select(0, slot.long())
    xt = xT[:, tmax].to(torch.float32).contiguous().clone()                             # [K, C, H, W]
    zs = zs.contiguous()
    ts_cpu = sched.timesteps_cpu[-n:]
    lane = (graph_lane if graph_lane is not None else 1) if (USE_CUDA_GRAPHS and xT.is_cuda) else 0
    x_in = torch.empty((K * rows, *xt.shape[1:]), device=model.device, dtype=torch.float32)
    for k_step in range(n):
        t = int(ts_cpu[k_step])
        pos = N - n + k_step
        idx = n - k_step - 1
        x_in.copy_(xt.repeat_interleave(rows, 0))
        t_in = torch.full((K * rows,), t, dtype=torch.int64, device=model.device)
        eps = _unet_eval(model, x_in, t_in, text, slot, cl_b, slot_key=("revK", K, tuple(sl)), lane=lane)
        out = torch.empty_like(xt)
        for k in range(K):
            e = eps[k * rows:(k + 1) * rows]
            model.k_cfg_rev_step(pos, float(etas[idx]), e, e[1:], P, cfg_map, xt[k:k + 1], zs[k, idx], out[k:k + 1])
        xt = out
    return xt, zs


# ------------------------------------------------------------------------------------------------------------------
# General path: the reference's per-step structure (needed when h-space / skip-connection taps are requested).
# ------------------------------------------------------------------------------------------------------------------
def _forward_general(model, x0, etas, prog_bar, prompts, cfg_scales, num_inference_steps, cutoff_points,
                     numerical_fix, extract_h_space, extract_skipconns, duration, first_order):
    have_cond = len(prompts) > 1 or prompts[0] != ""
    if have_cond:
        text_hs, text_cl, text_mask = model.encode_text(prompts)
        batch_size = len(prompts)
        cfg_scales_tensor, _ = _build_cfg_maps(batch_size, x0.shape[1:], cfg_scales, cutoff_points, model.device,
                                               x0.dtype, prompts)
    un_hs, un_cl, un_mask = model.encode_text([""], negative=True)
    timesteps = model.model.scheduler.timesteps.to(model.device)
    variance_noise_shape = model.get_noise_shape(x0, num_inference_steps)
    if type(etas) in [int, float]:
        etas = [etas] * model.model.scheduler.num_inference_steps
    xts = model.sample_xts_from_x0(x0, num_inference_steps=num_inference_steps)
    zs = torch.zeros(size=variance_noise_shape, device=model.device)
    extra_info = [None] * len(zs)
    hspaces, skipconns = [], []
    t_to_idx = _t_to_idx(timesteps)
    xt = x0
    op = tqdm(timesteps) if prog_bar else timesteps
    model.setup_extra_inputs(xt, init_timestep=timesteps[0], audio_end_in_s=duration)
    for t in op:
        idx = num_inference_steps - t_to_idx[int(t)] - 1
        xt = xts[idx + 1][None]
        xt_inp = model.model.scheduler.scale_model_input(xt, t)
        with torch.no_grad():
            out, out_hspace, out_skipconns = model.unet_forward(
                xt_inp, timestep=t, encoder_hidden_states=un_hs, class_labels=un_cl, encoder_attention_mask=un_mask)
            if have_cond:
                cond_out, cond_out_hspace, cond_out_skipconns = model.unet_forward(
                    xt_inp.expand(len(prompts), -1, -1, -1), timestep=t, encoder_hidden_states=text_hs,
                    class_labels=text_cl, encoder_attention_mask=text_mask)
        if have_cond:
            noise_pred = out.sample + (cfg_scales_tensor * (cond_out.sample - out.sample.expand(batch_size, -1, -1, -1))
                                       ).sum(axis=0).unsqueeze(0)
            noise_h_space = out_hspace + cfg_scales[0] * (cond_out_hspace - out_hspace)
            if extract_skipconns:
                noise_skipconns = {k: [out_skipconns[k][j] + cfg_scales[0] * (cond_out_skipconns[k][j] - out_skipconns[k][j])
                                       for j in range(len(out_skipconns[k]))] for k in out_skipconns}
        else:
            noise_pred = out.sample
            noise_h_space = out_hspace
            if extract_skipconns:
                noise_skipconns = out_skipconns
        hspaces.append(noise_h_space)
        if extract_skipconns:
            skipconns.append(noise_skipconns)
        xtm1 = xts[idx][None]
        z, xtm1, extra = model.get_zs_from_xts(xt, xtm1, noise_pred, t, eta=etas[idx], numerical_fix=numerical_fix,
                                               first_order=first_order)
        zs[idx] = z
        xts[idx] = xtm1
        extra_info[idx] = extra
    if zs is not None:
        zs[0] = torch.zeros_like(zs[0])
    if extract_h_space:
        return xt, zs, xts, extra_info, torch.concat(hspaces, axis=0)
    return xt, zs, xts, extra_info, torch.concat(hspaces, axis=0), skipconns


def _reverse_general(model, xT, tstart, fix_alpha, etas, prompts, neg_prompts, cfg_scales, prog_bar, zs, cutoff_points,
                     hspace_add, hspace_replace, skipconns_replace, zero_out_resconns, extract_h_space,
                     extract_skipconns, duration, first_order, extra_info):
    batch_size = len(prompts)
    text_hs, text_cl, text_mask = model.encode_text(prompts)
    un_hs, un_cl, un_mask = model.encode_text(neg_prompts, negative=True)
    cfg_scales_tensor, masks = _build_cfg_maps(batch_size, xT.shape[1:], cfg_scales, cutoff_points, model.device,
                                               xT.dtype, masks_too=True)
    xt = xT[tstart.max()].unsqueeze(0)
    if etas is None:
        etas = 0
    if type(etas) in [int, float]:
        etas = [etas] * model.model.scheduler.num_inference_steps
    assert len(etas) == model.model.scheduler.num_inference_steps
    timesteps = model.model.scheduler.timesteps.to(model.device)
    op = tqdm(timesteps[-zs.shape[0]:]) if prog_bar else timesteps[-zs.shape[0]:]
    t_to_idx = _t_to_idx(timesteps[-zs.shape[0]:])
    hspaces, skipconns = [], []
    model.setup_extra_inputs(xt, extra_info=extra_info, init_timestep=timesteps[-zs.shape[0]], audio_end_in_s=duration)
    N = model.model.scheduler.num_inference_steps
    for it, t in enumerate(op):
        idx = N - t_to_idx[int(t)] - (N - zs.shape[0] + 1)
        xt_inp = model.model.scheduler.scale_model_input(xt, t)

        def tap_args(weight):
            return dict(
                mid_block_additional_residual=(None if hspace_add is None else weight *
                                               (hspace_add[-zs.shape[0]:][it] if hspace_add.shape[0] > 1 else hspace_add)),
                replace_h_space=(None if hspace_replace is None else
                                 (hspace_replace[-zs.shape[0]:][it].unsqueeze(0) if hspace_replace.shape[0] > 1
                                  else hspace_replace)),
                zero_out_resconns=zero_out_resconns,
                replace_skip_conns=(None if skipconns_replace is None else
                                    (skipconns_replace[-zs.shape[0]:][it] if len(skipconns_replace) > 1
                                     else skipconns_replace)))
        with torch.no_grad():
            uncond_out, out_hspace, out_skipconns = model.unet_forward(
                xt_inp, timestep=t, encoder_hidden_states=un_hs, class_labels=un_cl, encoder_attention_mask=un_mask,
                **tap_args(None if hspace_add is None else 1 / (cfg_scales[0] + 1)))
        if prompts:
            with torch.no_grad():
                cond_out, cond_out_hspace, cond_out_skipconns = model.unet_forward(
                    xt_inp.expand(batch_size, -1, -1, -1), timestep=t, encoder_hidden_states=text_hs,
                    class_labels=text_cl, encoder_attention_mask=text_mask,
                    **tap_args(None if hspace_add is None else cfg_scales[0] / (cfg_scales[0] + 1)))
        z = zs[idx] if zs is not None else None
        z = z.unsqueeze(0)
        if prompts:
            noise_pred = uncond_out.sample + (cfg_scales_tensor * (cond_out.sample -
                                                                   uncond_out.sample.expand(batch_size, -1, -1, -1))
                                              ).sum(axis=0).unsqueeze(0)
            if extract_h_space or extract_skipconns:
                noise_h_space = out_hspace + cfg_scales[0] * (cond_out_hspace - out_hspace)
            if extract_skipconns:
                noise_skipconns = {k: [out_skipconns[k][j] + cfg_scales[0] * (cond_out_skipconns[k][j] - out_skipconns[k][j])
                                       for j in range(len(out_skipconns[k]))] for k in out_skipconns}
        else:
            noise_pred = uncond_out.sample
            if extract_h_space or extract_skipconns:
                noise_h_space = out_hspace
            if extract_skipconns:
                noise_skipconns = out_skipconns
        if extract_h_space or extract_skipconns:
            hspaces.append(noise_h_space)
        if extract_skipconns:
            skipconns.append(noise_skipconns)
        xt = model.reverse_step_with_custom_noise(noise_pred, t, xt, variance_noise=z, eta=etas[idx],
                                                  first_order=first_order)
        apply_fix = ((tstart.max() - tstart) > it)
        if apply_fix.any():
            apply_fix = (apply_fix * fix_alpha).unsqueeze(1).unsqueeze(2).unsqueeze(3).to(xT.device)
            xt = (masks * (xt.expand(batch_size, -1, -1, -1) * (1 - apply_fix) +
                           apply_fix * (xT[tstart.max() - it - 1].expand(batch_size, -1, -1, -1)))).sum(axis=0).unsqueeze(0)
    if extract_h_space:
        return xt, zs, torch.concat(hspaces, axis=0)
    if extract_skipconns:
        return xt, zs, torch.concat(hspaces, axis=0), skipconns
    return xt, zs
