"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Makes the *unmodified* reference (kabachuha/sd-webui-text2video) importable in a
container that has neither the Auto1111 webui (`modules.*`) nor Stability-AI's
`ldm.*` package.  Used only by `oracle/make_golden.py` (fixture generation) and by
the CPU tests that pin the oracle restatement against the real reference when
`/root/reference` exists.  Nothing under `-m gpu`, `smoke()` or `bench.py` imports this.

What is stubbed:
  * webui `modules.shared / prompt_parser / sd_samplers_common / sd_hijack_optimizations /
    paths / extensions / devices` -> inert ModuleType stubs (state flags, empty cmd_opts,
    identity reconstruct_cond_batch).
  * `ldm.*` (un-vendored third-party package "stablediffusion", unpinned by the reference)
    -> backed by the reference's OWN vendored twin of the same upstream code under
    scripts/videocrafter/lvdm (util.py:13-88, autoencoder_modules.py:382-596,
    distributions.py:5-46).  `noise_like` differs (ldm's has no generator arg) and is
    restated here with ldm's signature.
  * `omegaconf.listconfig.ListConfig` -> list.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("T2V_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "scripts", "modelscope"))


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Idempotently install the stub modules and put the reference's `scripts/` on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    import torch

    scripts = os.path.join(REF_ROOT, "scripts")
    if scripts not in sys.path:
        sys.path.insert(0, scripts)

    # ---- webui stubs -------------------------------------------------------------
    modules = _mod("modules")
    modules.__path__ = []
    shared = _mod("modules.shared")

    class _State:
        interrupted = False
        skipped = False
        sampling_step = 0
        sampling_steps = 0
        job = ""
        job_no = 0
        job_count = 0

    class _Opts:
        data = {}

        def __getattr__(self, k):
            return None

    shared.state = _State()
    shared.opts = _Opts()
    shared.cmd_opts = types.SimpleNamespace()  # empty -> opt_sdp_attention default True (t2v_model.py:566)
    shared.device = torch.device("cpu")
    shared.xformers_available = False
    modules.shared = shared

    pp = _mod("modules.prompt_parser")
    pp.reconstruct_cond_batch = lambda c, step: c
    modules.prompt_parser = pp

    sc = _mod("modules.sd_samplers_common")

    class InterruptedException(BaseException):
        pass

    sc.InterruptedException = InterruptedException
    modules.sd_samplers_common = sc

    hj = _mod("modules.sd_hijack_optimizations")
    hj.get_xformers_flash_attention_op = lambda q, k, v: None
    modules.sd_hijack_optimizations = hj

    paths = _mod("modules.paths")
    paths.models_path = "/nonexistent/models"
    modules.paths = paths
    ext = _mod("modules.extensions")
    ext.extensions = []
    modules.extensions = ext
    dev = _mod("modules.devices")
    dev.has_mps = lambda: False
    modules.devices = dev

    # ---- omegaconf stub (videocrafter openaimodel3d.py:9) ---------------------------
    if "omegaconf" not in sys.modules:
        oc = _mod("omegaconf")
        oc.__path__ = []
        lc = _mod("omegaconf.listconfig")
        lc.ListConfig = list
        oc.listconfig = lc
        oc.ListConfig = list

    # ---- ldm stubs backed by the reference's vendored twin ------------------------------
    from videocrafter.lvdm.models.modules import util as vc_util
    from videocrafter.lvdm.models.modules import autoencoder_modules as vc_ae
    from videocrafter.lvdm.models.modules import distributions as vc_dist
    from videocrafter.lvdm.utils import common_utils as vc_common

    ldm = _mod("ldm")
    ldm.__path__ = []
    lutil = _mod("ldm.util")
    lutil.instantiate_from_config = vc_common.instantiate_from_config
    _mod("ldm.modules").__path__ = []
    _mod("ldm.modules.diffusionmodules").__path__ = []
    du = _mod("ldm.modules.diffusionmodules.util")
    du.make_beta_schedule = vc_util.make_beta_schedule
    du.make_ddim_timesteps = vc_util.make_ddim_timesteps
    du.make_ddim_sampling_parameters = vc_util.make_ddim_sampling_parameters
    du.extract_into_tensor = vc_util.extract_into_tensor

    def noise_like(shape, device, repeat=False):
        # ldm signature (no generator); global RNG, as in Stability-AI/stablediffusion
        if repeat:
            return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
        return torch.randn(shape, device=device)

    du.noise_like = noise_like
    dm = _mod("ldm.modules.diffusionmodules.model")
    dm.Encoder = vc_ae.Encoder
    dm.Decoder = vc_ae.Decoder
    _mod("ldm.modules.distributions").__path__ = []
    dd = _mod("ldm.modules.distributions.distributions")
    dd.DiagonalGaussianDistribution = vc_dist.DiagonalGaussianDistribution
    _installed = True


def load_modelscope():
    """Returns the reference's modelscope/t2v_model module."""
    install()
    import importlib
    return importlib.import_module("modelscope.t2v_model")


def load_samplers():
    install()
    import importlib
    load_modelscope()
    return importlib.import_module("samplers.samplers_common")


def load_key_frames():
    """The reference's t2v_helpers/key_frames.py with `numexpr` (not installed here) replaced by a stand-in whose
    `evaluate(expr)` resolves names in the CALLER's frame, which is how numexpr finds `t`, `max_f`, `max_i_f`, `s`
    (key_frames.py:30-40, :84-88).  pandas is present and used unmodified."""
    install()
    import importlib
    import math
    import pandas                                            # noqa: F401  imported BEFORE the stand-in exists: pandas probes
    import pandas.core.computation.expressions               # noqa: F401  for an optional numexpr at import time
    if "numexpr" not in sys.modules:
        nx = _mod("numexpr")

        def evaluate(expr):
            f = sys._getframe(1)
            env = {k: getattr(math, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "floor", "ceil")}
            env.update({k: v for k, v in f.f_locals.items() if isinstance(v, (int, float))})
            return eval(expr, {"__builtins__": {}}, env)          # test infrastructure only: evaluates OUR fixed test strings
        nx.evaluate = evaluate
    return importlib.import_module("t2v_helpers.key_frames")
