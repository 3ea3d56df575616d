"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (HilaManor/AudioEditingCode) from
/root/reference so the restatements under oracle/ can be validated against it and golden vectors
generated (oracle/make_golden.py).  Works only in the build container (the GPU box has no
/root/reference); nothing in the product package, bench.py or `-m gpu` tests may import this file.

The reference's third-party imports that are absent here (diffusers, librosa, wandb, soundfile,
progressbar, omegaconf — SURVEY.md §8c) are replaced by inert stubs *before* the reference modules
are imported; no reference source is copied or edited.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("AEDIT_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_CODE, "models.py"))


class _Anything:
    """Placeholder class for names the reference imports but never touches on our test paths."""

    def __init__(self, *a, **k):
        raise RuntimeError("stubbed third-party class instantiated in the oracle harness")


def _mod(name, **attrs):
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class UNet2DConditionOutput:
    """3-line stand-in for diffusers.models.unets.unet_2d_condition.UNet2DConditionOutput
    (models.py:7,393 only constructs it with sample= and reads .sample)."""

    def __init__(self, sample=None):
        self.sample = sample


def _install_stubs():
    if "diffusers" not in sys.modules:
        names = ["DDIMScheduler", "UNet2DModel", "VQModel", "CosineDPMSolverMultistepScheduler",
                 "AudioLDMPipeline", "AudioLDM2Pipeline", "StableDiffusionPipeline", "StableAudioPipeline"]
        d = _mod("diffusers", **{n: type(n, (_Anything,), {}) for n in names})
        d.__path__ = []
        _mod("diffusers.schedulers").__path__ = []
        _mod("diffusers.schedulers.scheduling_dpmsolver_sde",
             BrownianTreeNoiseSampler=type("BrownianTreeNoiseSampler", (_Anything,), {}))
        _mod("diffusers.models").__path__ = []
        _mod("diffusers.models.unets").__path__ = []
        _mod("diffusers.models.unets.unet_2d_condition", UNet2DConditionOutput=UNet2DConditionOutput)
        _mod("diffusers.models.embeddings", get_1d_rotary_pos_embed=lambda *a, **k: None)
    for name in ("wandb", "soundfile", "progressbar"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _mod(name)
    if "omegaconf" not in sys.modules:
        try:
            importlib.import_module("omegaconf")
        except Exception:
            _mod("omegaconf").__path__ = []
            _mod("omegaconf.listconfig", ListConfig=type("ListConfig", (list,), {}))
    if "librosa" not in sys.modules:
        try:
            importlib.import_module("librosa")
        except Exception:
            _install_librosa_stub()
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib")
        except Exception:
            _mod("matplotlib").__path__ = []
            _mod("matplotlib.pyplot")


def slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(htk=False, norm='slaney') restated (librosa 0.9 algorithm, float32 output).
    Used (a) as the stub behind audioldm/audio/stft.py:5-6,141-143 and (b) by the product front end
    (which carries its own copy) — the two are compared in tests."""
    import numpy as np

    def hz_to_mel(f):
        f = np.asanyarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asanyarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        freqs = f_sp * m
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)

    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def _install_librosa_stub():
    import numpy as np

    def pad_center(data, size, axis=-1, **kwargs):
        n = data.shape[axis]
        lpad = int((size - n) // 2)
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (lpad, int(size - n - lpad))
        return np.pad(data, lengths, **kwargs)

    def tiny(x):
        return np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating) else np.float32).tiny

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
        return slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax if fmax is not None else sr / 2)

    lib = _mod("librosa")
    lib.__path__ = []
    _mod("librosa.util", pad_center=pad_center, tiny=tiny)
    _mod("librosa.filters", mel=mel)
    lib.util = sys.modules["librosa.util"]
    lib.filters = sys.modules["librosa.filters"]


def _bare_package(name, path):
    """Register a package WITHOUT running its __init__ (audioldm/__init__.py pulls librosa-heavy
    pipeline code we do not need)."""
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with the reference modules: models, inversion_utils, pc_drift,
    ddim_inversion, openaimodel, attention, util, vae_modules, hifigan, stft."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT} (only present in the build container)")
    import transformers  # noqa: F401  (real package; import before any stub shadows its probes)
    from transformers import RobertaTokenizer, RobertaTokenizerFast  # noqa: F401
    _install_stubs()
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    al = os.path.join(REF_CODE, "audioldm")
    _bare_package("audioldm", al)
    for sub in ("latent_diffusion", "variational_autoencoder", "hifigan", "audio"):
        _bare_package(f"audioldm.{sub}", os.path.join(al, sub))
    _loaded["audioldm_utils"] = importlib.import_module("audioldm.utils")
    _loaded["util"] = importlib.import_module("audioldm.latent_diffusion.util")
    _loaded["attention"] = importlib.import_module("audioldm.latent_diffusion.attention")
    _loaded["openaimodel"] = importlib.import_module("audioldm.latent_diffusion.openaimodel")
    _loaded["vae_modules"] = importlib.import_module("audioldm.variational_autoencoder.modules")
    _loaded["hifigan"] = importlib.import_module("audioldm.hifigan.models")
    _loaded["stft"] = importlib.import_module("audioldm.audio.stft")
    _loaded["audio_tools"] = importlib.import_module("audioldm.audio.tools")
    for name in ("models", "pc_drift", "utils", "ddm_inversion", "ddm_inversion.inversion_utils",
                 "ddm_inversion.ddim_inversion"):
        mod = sys.modules.get(name)       # a same-named module from elsewhere (e.g. the dropin/ shims) must not shadow
        if mod is not None and not str(getattr(mod, "__file__", "") or "").startswith(REF_CODE):
            del sys.modules[name]
    _loaded["models"] = importlib.import_module("models")
    _bare_package("ddm_inversion", os.path.join(REF_CODE, "ddm_inversion"))
    _loaded["inversion_utils"] = importlib.import_module("ddm_inversion.inversion_utils")
    _loaded["ddim_inversion"] = importlib.import_module("ddm_inversion.ddim_inversion")
    _loaded["pc_drift"] = importlib.import_module("pc_drift")
    return types.SimpleNamespace(**_loaded)
