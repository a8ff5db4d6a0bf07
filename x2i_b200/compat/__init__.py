"""``diffusers`` import shim: lets the UNMODIFIED X2I entry points resolve their ``diffusers`` imports to the x2i_b200 drop-ins.

The reference's hot path lives behind ``import diffusers`` (``train/train_qwenvl.py:29-47``, ``infer/inference_qwenvl.py:7-8``,
``lightcontrol/lightcontrol_flux.py:22-39``, ``lightcontrol/train_lightcontrol.py:27-38``); ``diffusers==0.31.0`` is neither vendored
nor installable here.  ``install()`` registers a synthetic ``diffusers`` package in ``sys.modules`` whose names are the same-named
x2i_b200 classes, so

    import x2i_b200.compat as compat; compat.install()
    import train_qwenvl            # the reference file, unchanged: FluxTransformer2DModel is x2i_b200.flux.FluxTransformer2DModel

Names of diffusers that are outside the hot path (LoRA mixins, hub utilities, EMA, wandb probes) are present as inert placeholders so
the reference modules import; calling one raises ``X2IError``.  Small host-side utilities the train-step skeleton actually uses are
real: ``optimization.get_scheduler`` (cosine / linear / constant with warm-up, ``train_qwenvl.py:476-481``),
``training_utils.compute_density_for_timestep_sampling`` / ``compute_loss_weighting_for_sd3`` (``train_lightcontrol.py:690-696,:753``),
``utils.torch_utils.randn_tensor`` / ``is_compiled_module``.
"""
from __future__ import annotations

import math
import sys
import types

import torch

from .._lib import X2IError

__all__ = ["install", "uninstall", "get_scheduler"]
_INSTALLED = []


def _inert(name):
    def _raise(*a, **k):
        raise X2IError(f"diffusers.{name} is outside the X2I hot path and not provided by x2i_b200.compat")

    class _Inert:  # usable as a base class / mixin and importable by name; instantiating the bare placeholder raises
        def __init__(self, *a, **k):
            if type(self) is _Inert:
                _raise()

    _Inert.__name__ = _Inert.__qualname__ = name.rsplit(".", 1)[-1]
    return _Inert


# ---------------------------------------------------------------------------------------------- diffusers.optimization
def get_scheduler(name, optimizer, num_warmup_steps=None, num_training_steps=None, num_cycles: float = 0.5, power: float = 1.0, **kw):
    """``diffusers.optimization.get_scheduler`` [D031] for the schedules the reference scripts select (``--lr_scheduler``,
    train_qwenvl.py:476-481: "cosine" with 100 warm-up steps in train_qwenvl.sh): returns a ``LambdaLR``."""
    from torch.optim.lr_scheduler import LambdaLR
    name = getattr(name, "value", name)
    opt = getattr(optimizer, "optimizer", optimizer)  # MasterWeightOptimizer wraps the real one
    w = int(num_warmup_steps or 0)
    n = num_training_steps

    def warm(step):
        return float(step) / float(max(1, w))

    if name == "constant":
        return LambdaLR(opt, lambda _: 1.0)
    if name == "constant_with_warmup":
        return LambdaLR(opt, lambda s: warm(s) if s < w else 1.0)
    if n is None:
        raise ValueError(f"{name} requires `num_training_steps`, please provide that argument.")
    if name == "linear":
        return LambdaLR(opt, lambda s: warm(s) if s < w else max(0.0, float(n - s) / float(max(1, n - w))))
    if name == "cosine":
        def f(s):
            if s < w:
                return warm(s)
            progress = float(s - w) / float(max(1, n - w))
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))
        return LambdaLR(opt, f)
    if name == "polynomial":
        lr_init = opt.defaults["lr"]
        lr_end = 1e-7

        def f(s):
            if s < w:
                return warm(s)
            if s > n:
                return lr_end / lr_init
            return ((lr_init - lr_end) * (1 - (s - w) / (n - w)) ** power + lr_end) / lr_init
        return LambdaLR(opt, f)
    raise X2IError(f"get_scheduler: schedule '{name}' is not provided by x2i_b200.compat")


def compute_loss_weighting_for_sd3(weighting_scheme, sigmas=None):
    """diffusers.training_utils [D031] (train_lightcontrol.py:753): "sigma_sqrt" -> sigma^-2, "cosmap" -> 2 / (pi (1 - 2s + 2s^2)), else 1."""
    if weighting_scheme == "sigma_sqrt":
        return (sigmas ** -2.0).float()
    if weighting_scheme == "cosmap":
        return 2 / (math.pi * (1 - 2 * sigmas + 2 * sigmas ** 2))
    return torch.ones_like(sigmas)


def _is_compiled_module(module):
    return hasattr(torch, "_dynamo") and isinstance(module, torch._dynamo.eval_frame.OptimizedModule)


def _build():
    from .. import controlnext, flux, pipeline, vae
    from ..train_lightcontrol import compute_density_for_timestep_sampling
    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__x2i_compat__ = True
        mods[name] = m
        return m

    class Transformer2DModelOutput(types.SimpleNamespace):
        pass

    class BaseOutput(dict):
        pass

    def register_to_config(init):
        return init

    def maybe_allow_in_graph(cls):
        return cls

    def is_torch_version(op, version):
        from packaging import version as V
        import operator
        ops_ = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "<": operator.lt, "<=": operator.le, "!=": operator.ne}
        return ops_[op](V.parse(V.parse(torch.__version__).base_version), V.parse(version))

    import logging as _pylog

    class _Logging(types.SimpleNamespace):
        @staticmethod
        def get_logger(name=None):
            return _pylog.getLogger(name)

        @staticmethod
        def set_verbosity_info():
            return None

        set_verbosity_warning = set_verbosity_error = set_verbosity_info

    logging = _Logging()

    mod("diffusers", __version__="0.31.0", __path__=[], FluxPipeline=pipeline.FluxPipeline, AutoencoderKL=vae.AutoencoderKL,
        FluxTransformer2DModel=flux.FluxTransformer2DModel, FlowMatchEulerDiscreteScheduler=pipeline.FlowMatchEulerDiscreteScheduler,
        FluxControlNetPipeline=_inert("FluxControlNetPipeline"), FluxControlNetModel=_inert("FluxControlNetModel"))
    mod("diffusers.image_processor", VaeImageProcessor=vae.VaeImageProcessor)
    mod("diffusers.schedulers", __path__=[], FlowMatchEulerDiscreteScheduler=pipeline.FlowMatchEulerDiscreteScheduler)
    mod("diffusers.pipelines", __path__=[], FluxPipeline=pipeline.FluxPipeline)
    mod("diffusers.models", __path__=[], FluxTransformer2DModel=flux.FluxTransformer2DModel, AutoencoderKL=vae.AutoencoderKL)
    mod("diffusers.models.transformers", __path__=[], FluxTransformer2DModel=flux.FluxTransformer2DModel)
    mod("diffusers.models.attention", FeedForward=flux.FeedForward)
    mod("diffusers.models.attention_processor", Attention=flux.Attention, AttentionProcessor=flux.AttentionProcessor,
        FluxAttnProcessor2_0=flux.FluxAttnProcessor2_0, FusedFluxAttnProcessor2_0=flux.FusedFluxAttnProcessor2_0)
    mod("diffusers.models.normalization", AdaLayerNormContinuous=flux.AdaLayerNormContinuous, AdaLayerNormZero=flux.AdaLayerNormZero,
        AdaLayerNormZeroSingle=flux.AdaLayerNormZeroSingle, RMSNorm=flux.RMSNorm)
    mod("diffusers.models.embeddings", CombinedTimestepGuidanceTextProjEmbeddings=flux.CombinedTimestepGuidanceTextProjEmbeddings,
        CombinedTimestepTextProjEmbeddings=flux.CombinedTimestepTextProjEmbeddings, FluxPosEmbed=flux.FluxPosEmbed,
        TimestepEmbedding=controlnext.TimestepEmbedding, Timesteps=_inert("models.embeddings.Timesteps"))
    mod("diffusers.models.modeling_outputs", Transformer2DModelOutput=Transformer2DModelOutput)
    mod("diffusers.models.modeling_utils", ModelMixin=torch.nn.Module)
    mod("diffusers.models.resnet", Downsample2D=controlnext.Downsample2D, ResnetBlock2D=controlnext.ResnetBlock2D)
    mod("diffusers.configuration_utils", ConfigMixin=object, register_to_config=register_to_config)
    mod("diffusers.loaders", __path__=[], FromOriginalModelMixin=object, PeftAdapterMixin=object)
    mod("diffusers.loaders.lora_pipeline", SD3LoraLoaderMixin=_inert("SD3LoraLoaderMixin"))
    mod("diffusers.utils", __path__=[], USE_PEFT_BACKEND=False, is_torch_version=is_torch_version, logging=logging, BaseOutput=BaseOutput,
        scale_lora_layers=lambda *a, **k: None, unscale_lora_layers=lambda *a, **k: None, check_min_version=lambda v: None,
        is_wandb_available=lambda: False, get_peft_kwargs=_inert("utils.get_peft_kwargs"), get_adapter_name=_inert("utils.get_adapter_name"))
    mod("diffusers.utils.torch_utils", maybe_allow_in_graph=maybe_allow_in_graph, is_compiled_module=_is_compiled_module,
        randn_tensor=pipeline.randn_tensor)
    mod("diffusers.utils.hub_utils", load_or_create_model_card=_inert("utils.hub_utils.load_or_create_model_card"),
        populate_model_card=_inert("utils.hub_utils.populate_model_card"))
    mod("diffusers.optimization", get_scheduler=get_scheduler)
    mod("diffusers.training_utils", EMAModel=_inert("EMAModel"), compute_snr=_inert("compute_snr"),
        compute_density_for_timestep_sampling=compute_density_for_timestep_sampling,
        compute_loss_weighting_for_sd3=compute_loss_weighting_for_sd3)
    # wire the parent -> child attributes (``import diffusers; diffusers.utils.check_min_version(...)``)
    for name, m in mods.items():
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(mods[parent], child, m)
    return mods


def install(force: bool = False):
    """Register the shim under ``diffusers``.  A real diffusers install wins unless ``force`` is set."""
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "__x2i_compat__", False) and not force:
        return sys.modules["diffusers"]
    if not force and "diffusers" not in sys.modules:
        import importlib.util
        try:
            if importlib.util.find_spec("diffusers") is not None:
                import diffusers  # noqa: F401  (the genuine package is present: leave it alone)
                return sys.modules["diffusers"]
        except (ImportError, ValueError):
            pass
    mods = _build()
    sys.modules.update(mods)
    _INSTALLED[:] = list(mods)
    return mods["diffusers"]


def uninstall():
    for name in _INSTALLED:
        if getattr(sys.modules.get(name), "__x2i_compat__", False):
            sys.modules.pop(name, None)
    _INSTALLED.clear()
