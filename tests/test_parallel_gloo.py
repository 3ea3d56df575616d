"""N > 1 host logic on CPU: world_size-2 gloo process group (127.0.0.1 rendezvous) exercising clip sharding, the
ordered all-gather of results and the max-over-ranks timing reduction of audioeditingcode_b200.parallel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from audioeditingcode_b200 import parallel as P
    assert P.shard_indices(n_clips, rank, ws) == list(range(rank, n_clips, ws))
    clips = [torch.full((1, 8, 4, 16), float(i)) for i in range(n_clips)]
    calls = []

    def edit(x):
        calls.append(int(x.flatten()[0]))
        return x * 2 + 1
    out = P.edit_clips(edit, clips)
    ok = all(torch.equal(out[i], clips[i] * 2 + 1) for i in range(n_clips))
    ok = ok and calls == list(range(rank, n_clips, ws))
    local_only = P.edit_clips(edit, clips, gather=False)
    ok = ok and all((local_only[i] is not None) == (i % ws == rank) for i in range(n_clips))
    mx = P.max_over_ranks(10.0 + rank)
    ok = ok and mx == 10.0 + (ws - 1)
    # row-ownership merge used by the timestep-sharded forward process
    full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3) + 1
    mine = full.clone()
    owned = [i for i in range(7) if (i // 2) % ws == rank]
    mine[[i for i in range(7) if i not in owned]] = -99.0          # stale rows on this rank
    ok = ok and torch.equal(P.merge_owned_rows_(mine, owned), full)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 4, 1])
def test_clip_sharding_gloo_world2(n_clips):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_clips, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------ pc_drift: eigen-direction sharding + iterate all-reduce
class _FakeLDM:
    """Batch-row-independent stand-in for the wrapper protocol pc_drift needs (unet_forward / scheduler.step /
    get_sigma) so that the distributed result can be compared bit for bit with the single-process one on CPU."""

    def __init__(self):
        import types
        from oracle.ddpm_oracle import MiniDDIM
        sch = MiniDDIM(0.0015, 0.0195)
        sch.set_timesteps(20)
        self.model = types.SimpleNamespace(scheduler=sch)

    def unet_forward(self, x, timestep, encoder_hidden_states=None, class_labels=None, encoder_attention_mask=None):
        import types
        y = torch.tanh(x + 0.5 * torch.roll(x, 1, 2) - 0.25 * torch.roll(x, 1, 3))
        y = y * (1 + 0.1 * encoder_hidden_states.sum((1, 2)).view(-1, 1, 1, 1))
        return types.SimpleNamespace(sample=y), None, None

    def get_sigma(self, t):
        a = self.model.scheduler.alphas_cumprod[int(t)]
        return ((1 - a) / a) ** 0.5


def _pc_inputs(n_ev):
    from audioeditingcode_b200.pc_drift import PromptEmbeddings
    g = torch.Generator().manual_seed(11)
    xt = torch.randn(1, 8, 8, 16, generator=g)
    lat = torch.randn(1, 8, 8, 16, generator=g)
    mask = torch.ones(1, 8, 8, 16)
    mask[..., :2] = 0
    unc = PromptEmbeddings(torch.randn(1, 4, 6, generator=g), None, None)
    txt = PromptEmbeddings(torch.randn(1, 4, 6, generator=g), None, None)
    return xt, lat, mask, unc, txt


def _pc_worker(rank, ws, port, n_ev, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from audioeditingcode_b200 import pc_drift as PC
    ldm = _FakeLDM()
    xt, lat, mask, unc, txt = _pc_inputs(n_ev)
    t = ldm.model.scheduler.timesteps[5]
    x0_pred = PC.forward_directional(ldm, xt, t, lat, unc, txt, 3.0, eta=1, eigvecs=0, amount=0)[1] * mask
    kw = dict(pc_mode=PC.PCStreamChoice.BOTH, const=1e-3, cfg_tar=3.0, iters=6, eta=1, n_ev=n_ev)
    torch.manual_seed(0)
    ref = PC.get_eigenvectors(ldm, xt, txt, unc, lat, mask, t, x0_pred, **kw)          # single process
    torch.manual_seed(0 if rank == 0 else 123)                                          # start comes from rank 0
    got = PC.get_eigenvectors(ldm, xt, txt, unc, lat, mask, t, x0_pred, group=dist.group.WORLD, **kw)
    ok = torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
    ok = ok and all(torch.equal(a, b) for a, b in zip(ref[3], got[3]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_ev", [3, 2, 1])
def test_pc_drift_eigvec_sharding_gloo_world2(n_ev):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pc_worker, args=(r, ws, port, n_ev, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
