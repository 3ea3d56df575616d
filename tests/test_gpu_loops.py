"""GPU parity of the scheduler kernels and of the inversion loops through the drop-in wrapper API.
  * scheduler kernels (ae_sample_xts / ae_cfg_inv_step / ae_cfg_rev_step): BIT-EXACT against the golden tensors
    produced by the unmodified reference when fed the reference's own noise predictions;
  * loops with the CUDA U-Net: tolerance vs the reference's golden zs / xts / edited latent, and the
    weight-independent replay invariant bit-exact in step-sequential mode."""
import pytest
import torch

from oracle import ddpm_oracle as D
from oracle import unet_torch as U
from tests.helpers import load_golden, tiny_cfg_and_weights, oracle_unet_fn, make_sched

pytestmark = pytest.mark.gpu


def _wrapper(n_steps, pred="epsilon", name="tiny-audioldm"):
    from audioeditingcode_b200 import models, unet_config as C
    import dataclasses
    cfg = dataclasses.replace(C.preset(name), prediction_type=pred)
    w = U.synthetic_weights(cfg, seed=0)
    m = models.load_model("synthetic/audioldm-tiny", torch.device("cuda"), n_steps, weights=w, config=cfg)
    return m


class _GoldText:
    """encode_text stub returning the golden class-label vectors (the golden run used seeded vectors)."""

    def __init__(self, g, which):
        self.g, self.which = g, which

    def __call__(self, prompts, **kw):
        if prompts == [""]:
            return None, self.g["uncond"].cuda(), None
        return None, self.g[self.which].cuda(), None


@pytest.mark.parametrize("name,pred", [("loop_eps_single.npz", "epsilon"), ("loop_vpred_single.npz", "v_prediction")])
def test_sched_kernels_bitexact_vs_reference(name, pred):
    """Drive the wrapper's a3/a4/a5 methods with the ORACLE's fp32 noise predictions: every tensor must equal the
    reference's golden tensors bit for bit (only the scheduler kernels are on the GPU here)."""
    g = load_golden(name)
    cfg, w = tiny_cfg_and_weights()
    N = int(g["n_steps"])
    m = _wrapper(N, pred)
    sched = make_sched(cfg, N, pred)
    xts = m.sample_xts_from_x0(g["x0"].cuda(), N, noise=g["noise"].cuda())
    ref_xts = D.sample_xts_from_x0(sched, g["x0"], g["noise"])
    assert torch.equal(xts.cpu(), ref_xts)
    fn = oracle_unet_fn(cfg, w, g["uncond"], g["src"])
    cfgm, _ = D.build_cfg_maps(1, g["x0"].shape[1:], [float(g["cfg_src"][0])], None)
    zs = torch.zeros(N, *g["x0"].shape[1:])
    xts_c = xts.cpu().clone()
    for t in sched.timesteps:
        idx = N - int((sched.timesteps == t).nonzero()) - 1
        xt = xts_c[idx + 1][None]
        eps = D.cfg_combine(fn(xt, int(t), "uncond"), fn(xt, int(t), "cond"), cfgm)
        z, xtm1, _ = m.get_zs_from_xts(xt.cuda(), xts_c[idx][None].cuda(), eps.cuda(), t, eta=1.0, numerical_fix=True)
        zs[idx] = z.cpu()[0]
        xts_c[idx] = xtm1.cpu()[0]
    zs[0] = 0
    assert torch.equal(zs, g["zs"])
    assert torch.equal(xts_c, g["xts"])
    # reverse step kernel
    tstart = int(g["tstart"][0])
    fn_t = oracle_unet_fn(cfg, w, g["uncond"], g["tgt"])
    cfgt, _ = D.build_cfg_maps(1, g["x0"].shape[1:], [float(g["cfg_tar"][0])], None)
    xt = g["xts"][tstart][None]
    for k, t in enumerate(sched.timesteps[-tstart:]):
        idx = tstart - k - 1
        eps = D.cfg_combine(fn_t(xt, int(t), "uncond"), fn_t(xt, int(t), "cond"), cfgt)
        xt = m.reverse_step_with_custom_noise(eps.cuda(), t, xt.cuda(), variance_noise=g["zs"][idx][None].cuda(),
                                              eta=1.0).cpu()
    assert torch.equal(xt, g["w_edit"])


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("name,pred,fb", [("loop_eps_single.npz", "epsilon", 1), ("loop_eps_single.npz", "epsilon", 4),
                                          ("loop_vpred_single.npz", "v_prediction", 1),
                                          ("loop_eps_multi.npz", "epsilon", 1)])
def test_loops_vs_reference_golden(name, pred, fb):
    from audioeditingcode_b200.ddm_inversion import inversion_utils as IU
    g = load_golden(name)
    N = int(g["n_steps"])
    m = _wrapper(N, pred)
    m.encode_text = _GoldText(g, "src")
    P = g["src"].shape[0]
    xt, zs, xts, _ = IU.inversion_forward_process(
        m, g["x0"].cuda(), etas=1.0, prompts=["p%d" % i for i in range(P)], cfg_scales=[float(v) for v in g["cfg_src"]],
        num_inference_steps=N, numerical_fix=True, forward_batch=fb, noise=g["noise"].cuda())
    assert torch.count_nonzero(zs[0]) == 0
    # z divides a small residual by sigma_t: compare in units of the latent scale
    assert _rel(xts, g["xts"]) < 2e-3
    assert (zs.cpu() - g["zs"]).abs().max().item() < 0.35
    assert _rel(zs, g["zs"]) < 6e-2
    m.encode_text = _GoldText(g, "tgt")
    tstart = g["tstart"].to(torch.int)
    skip = N - tstart
    w_edit, _ = IU.inversion_reverse_process(
        m, xT=xts, tstart=tstart, etas=1.0, prompts=["q%d" % i for i in range(P)], neg_prompts=[""],
        cfg_scales=[float(v) for v in g["cfg_tar"]], zs=zs[:int(N - min(skip))])
    assert _rel(w_edit, g["w_edit"]) < 5e-2


def test_replay_invariant_bitexact_sequential():
    """F9 within our own stack: same prompt / cfg replays wts[k] bit-exactly for k >= 1 (forward_batch=1)."""
    from audioeditingcode_b200.ddm_inversion import inversion_utils as IU
    g = load_golden("loop_eps_single.npz")
    N = int(g["n_steps"])
    m = _wrapper(N)
    m.encode_text = _GoldText(g, "src")
    _, zs, xts, _ = IU.inversion_forward_process(m, g["x0"].cuda(), etas=1.0, prompts=["p"], cfg_scales=[3.0],
                                                 num_inference_steps=N, numerical_fix=True, forward_batch=1,
                                                 noise=g["noise"].cuda())
    finals = []
    for tstart in (N, N // 2):
        trace = []
        w, _ = IU.inversion_reverse_process(m, xT=xts, tstart=torch.tensor([tstart], dtype=torch.int), etas=1.0,
                                            prompts=["p"], neg_prompts=[""], cfg_scales=[3.0], zs=zs[:tstart],
                                            trace=trace)
        for k, xt in enumerate(trace):
            level = tstart - k - 1
            if level >= 1:
                assert torch.equal(xt[0], xts[level]), f"replay mismatch at level {level} (tstart={tstart})"
        finals.append(w)
    assert torch.equal(finals[0], finals[1])
