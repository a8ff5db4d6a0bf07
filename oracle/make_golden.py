"""Mint the golden fixtures under tests/golden/ from the REFERENCE ITSELF (run in the build
container, where /root/reference is mounted; the GPU box never runs this).

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  What each fixture pins, and how:

  proj_small.pt / proj_c1.pt   reference ``utils/proj.py`` imported unmodified; synthetic weights are
                               written into it by state-dict key; outputs stored.
  proj_legacy.pt               reference ``model_internvl/proj.py`` (a12) imported unmodified with the real ``transformers`` T5Stack: Proj,
                               Proj2, Proj3, MLP, MLP2, MLP_plus outputs on seeded inputs and weights.
  helpers.pt                   reference ``train/train_qwenvl.py`` imported with every missing third-party
                               module auto-stubbed; ``normalize``, ``_prepare_latent_image_ids``,
                               ``_pack_latents``, ``calculate_shift`` called directly.
  kd_loop.pt                   reference ``train/train_qwenvl.py`` imported with ``x2i_b200.compat`` as its ``diffusers``: its own
                               ``cast_hook_list`` run on the x2i_b200 transformer, and the LITERAL loss loop (:601-620, cut out of
                               ``train()`` with inspect and exec'ed) on seeded tensors incl. the inf/nan guard.
  flux_structure.pt            reference ``lightcontrol/lightcontrol_flux.py`` block / transformer classes
                               imported unmodified, with a fake ``diffusers`` package whose LEAF classes are
                               this repo's oracle leaves (diffusers 0.31.0 is absent).  Pins the block
                               wiring (order of modulation chunks, gates, residuals, cat order, slicing);
                               does NOT pin the leaves.
  resampler_small.pt           reference ``minicpm/resampler.py`` imported unmodified (with the 3 builtins it expects torch
                               to leak); ragged tgt_sizes; output + its sincos table stored.
  crosscheck.json              oracle blocks vs the BFL-derived Flux blocks shipped in ``torchtitan``
                               (independent implementation, weights remapped) -- max-abs differences.
                               Sanity evidence for the leaves, not a parity authority.
"""
import functools
import importlib
import importlib.abc
import importlib.machinery
import inspect
import json
import os
import sys
import types
from types import SimpleNamespace

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import flux_oracle as fo  # noqa: E402
from oracle import kd_oracle, proj_oracle  # noqa: E402


def synth_state(module: nn.Module, seed: int, std: float = 0.05):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in module.state_dict().items():
        if v.ndim >= 2:
            sd[k] = torch.randn(v.shape, generator=g) * std
        elif "norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = torch.randn(v.shape, generator=g) * std
    return sd


# ----------------------------------------------------------------------------- stubs
class _Meta(type):
    """Lets the stub CLASSES themselves be used as attribute bags / dicts (rpyc.core.protocol.X[...] = ...)."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Meta(name, (_Anything,), {})

    def __getitem__(cls, k):
        return _Anything()

    def __setitem__(cls, k, v):
        pass


class _Anything(metaclass=_Meta):
    """Class usable as base class, decorator, callable and attribute bag."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = _Meta(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Last-resort finder: any module nobody else can find becomes a stub."""

    # third-party roots the reference imports that are absent from this image
    ROOTS = ("diffusers", "accelerate", "webdataset", "pytorch_lightning", "deepspeed", "bitsandbytes", "timm",
             "decord", "librosa", "soundfile", "rpyc", "braceexpand", "qwen_vl_utils", "peft", "ray", "cv2",
             "torchvision", "datasets", "apex", "xformers", "wandb", "tensorboard", "lightning_fabric")

    def __init__(self):
        self.stubbed = []

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] not in self.ROOTS:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        self.stubbed.append(spec.name)
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ----------------------------------------------------------------------------- A: projector
def golden_projector():
    sys.path.insert(0, REF)
    ref_proj = importlib.import_module("utils.proj")  # the reference file, unmodified
    out = {}
    # small, every branch (cnn / scale / mean)
    for tag, kw in (("cnn", dict(use_scale=False, use_cnn=True)), ("scale", dict(use_scale=True, use_cnn=False)),
                    ("mean", dict(use_scale=False, use_cnn=False))):
        m = ref_proj.Proj7Exp(in_channels=5, kernel_size=5, input_dim=64, output_dim0=24, output_dim1=96,
                              use_t5=False, **kw)
        sd = synth_state(m, 11)
        m.load_state_dict(sd)
        x = torch.randn(2, 5, 9, 64, generator=torch.Generator().manual_seed(12))
        with torch.no_grad():
            pooled, seq = m(x)
        out[tag] = dict(state=sd, x=x, pooled=pooled, seq=seq)
    torch.save(out, os.path.join(OUT, "proj_small.pt"))
    # config 1 (BASELINE.json configs[0]): create_proj3_qwen3b(37, use_cnn), x[1,37,77,2048]; x and the
    # weights are regenerated from seeds by the test (too big to store), outputs stored.
    m = ref_proj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True)
    m.load_state_dict(synth_state(m, 21, std=0.02))
    x = torch.randn(1, 37, 77, 2048, generator=torch.Generator().manual_seed(0))
    t5 = torch.randn(1, 77, 4096, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        pooled, seq = m(x)
        mse = torch.nn.functional.mse_loss(seq, t5)
    torch.save(dict(pooled=pooled.clone(), seq_sub=seq[:, ::7, ::13].contiguous().clone(), seq_sum=float(seq.double().sum()), mse=float(mse),
                    n_params=sum(p.numel() for p in m.parameters())), os.path.join(OUT, "proj_c1.pt"))
    sys.path.remove(REF)
    print("projector goldens written; C1 mse =", float(mse))


# ----------------------------------------------------------------------------- B: train helpers
def golden_helpers():
    finder = _StubFinder()
    # nothing from transformers is needed by the helper functions; stub it too (its lazy importer probes
    # optional deps that are stubbed here and trips over them)
    finder.ROOTS = finder.ROOTS + ("transformers",)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] == "transformers"}
    sys.meta_path.insert(0, finder)  # stubs win over half-working installs (e.g. `datasets` here)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "train"))
    try:
        ref_train = importlib.import_module("train_qwenvl")
    finally:
        sys.meta_path.remove(finder)
        for k in [k for k in sys.modules if k.split(".")[0] in finder.ROOTS]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 7, 48, generator=g) * 2 + 0.3
    lat = torch.randn(2, 16, 8, 12, generator=g)
    out = dict(
        norm_in=x, norm_out=ref_train.normalize(x),
        ids_8x12=ref_train._prepare_latent_image_ids(2, 8, 12, "cpu", torch.float32),
        ids_128=ref_train._prepare_latent_image_ids(1, 128, 128, "cpu", torch.float32),
        pack_in=lat, pack_out=ref_train._pack_latents(lat, 2, 16, 8, 12),
        shift_4096=ref_train.calculate_shift(4096), shift_1024=ref_train.calculate_shift(1024),
        shift_dev_4096=ref_train.calculate_shift(4096, 256, 4096, 0.5, 1.15),
        stubbed=sorted(set(s.split(".")[0] for s in finder.stubbed)),
    )
    torch.save(out, os.path.join(OUT, "helpers.pt"))
    for p in (REF, os.path.join(REF, "train")):
        sys.path.remove(p)
    print("helper goldens written; stubbed third-party roots:", out["stubbed"])


# ----------------------------------------------------------------------------- B1b: the reference's own hook registration + loss loop
def golden_kd_loop():
    """``train/train_qwenvl.py`` imported with ``x2i_b200.compat`` standing in for ``diffusers`` (so the file's
    ``FluxTransformer2DModel`` IS the x2i_b200 drop-in) and every other absent third-party root stubbed.  Then
      (1) the reference's own ``cast_hook_list`` (:206-214) is run on an x2i_b200 transformer (meta device: registration needs no
          GPU) and the hook fan-out is recorded;
      (2) the LITERAL loss loop -- the source lines of ``train()`` from ``loss = 0`` (:601) to the line before ``train_loss = loss``
          (:622), cut out of the file with ``inspect`` and exec'ed unchanged -- runs on seeded tensors in fp32 and in bf16 (the
          reference's dtype), once on clean inputs and once with a NaN planted in one teacher layer (the inf/nan guard :606-609);
          losses and student gradients are stored.
    Pins oracle/kd_oracle.py's restated loop AND (in the -m gpu tests) x2i_b200.kd.attention_distillation_loss to the reference's
    own lines."""
    import contextlib
    import io
    import textwrap
    from x2i_b200 import compat
    from x2i_b200.flux import FluxTransformer2DModel
    finder = _StubFinder()
    finder.ROOTS = tuple(r for r in finder.ROOTS if r != "diffusers") + ("transformers",)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("transformers", "diffusers")}
    compat.install(force=True)
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "train"))
    try:
        ref_train = importlib.import_module("train_qwenvl")
        assert ref_train.FluxTransformer2DModel is FluxTransformer2DModel  # the drop-in is what the reference file now names
        with torch.device("meta"):
            model = FluxTransformer2DModel(num_layers=3, num_single_layers=4, num_attention_heads=2, joint_attention_dim=64,
                                           pooled_projection_dim=32)
        lists = []
        ref_train.cast_hook_list(model, lists)  # the reference's function, verbatim
        n_hooks = [len(b.attn._forward_hooks) for b in list(model.transformer_blocks) + list(model.single_transformer_blocks)]
        # fire the registered callbacks the way nn.Module.__call__ / FluxTrainFn do: (module, input, output)
        for i, b in enumerate(model.transformer_blocks):
            for h in b.attn._forward_hooks.values():
                h(b.attn, (), (torch.full((1,), float(i)), torch.full((1,), 100.0 + i)))
        for i, b in enumerate(model.single_transformer_blocks):
            for h in b.attn._forward_hooks.values():
                h(b.attn, (), torch.full((1,), 200.0 + i))
        fanout = [[float(t) for t in lst] for lst in lists]
        src = inspect.getsource(ref_train.train).splitlines()
        a = next(i for i, l in enumerate(src) if l.strip() == "loss = 0")
        b = next(i for i, l in enumerate(src) if l.strip() == "train_loss = loss")
        loop_src = textwrap.dedent("\n".join(src[a:b]))
        assert "for i in range(19):" in loop_src and "for i in range(38):" in loop_src and "batchmean" in loop_src
        normalize = ref_train.normalize
    finally:
        sys.meta_path.remove(finder)
        for p_ in (REF, os.path.join(REF, "train")):
            sys.path.remove(p_)
        compat.uninstall()
        for k in [k for k in sys.modules if k.split(".")[0] in finder.ROOTS]:
            sys.modules.pop(k)
        sys.modules.update(saved)

    g = torch.Generator().manual_seed(31)
    B, D = 2, 64
    shapes = ((B, 19, 6, D), (B, 19, 4, D), (B, 38, 8, D))
    teacher = [(torch.randn(s_, generator=g) * 1.5 + 0.2).bfloat16() for s_ in shapes]
    student = [(t.float() + 0.4 * torch.randn(t.shape, generator=g)).bfloat16() for t in teacher]

    def run(ts, ss, dtype):
        ss = [x.to(dtype).clone().requires_grad_(True) for x in ss]
        ns = dict(torch=torch, F=torch.nn.functional, normalize=normalize,
                  KD_teacher_tensor0=ts[0].to(dtype), KD_teacher_tensor1=ts[1].to(dtype), KD_teacher_tensor2=ts[2].to(dtype),
                  KD_student_tensor0=ss[0], KD_student_tensor1=ss[1], KD_student_tensor2=ss[2])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            exec(loop_src, ns)  # the reference's lines, unchanged
        loss = ns["loss"]
        grads = torch.autograd.grad(loss, ss)
        return dict(loss=loss.detach().float(), grads=[g_.float() for g_ in grads], printed=buf.getvalue().split())

    poisoned = [t.clone() for t in teacher]
    poisoned[2][0, 5, 3, 7] = float("nan")   # one element of single-block layer 5
    poisoned[0][1, 2, 0, 0] = float("inf")   # one element of double-block image layer 2
    out = dict(teacher=teacher, student=student, loop_src=loop_src, n_hooks=n_hooks, hook_fanout=fanout,
               clean_fp32=run(teacher, student, torch.float32), clean_bf16=run(teacher, student, torch.bfloat16),
               poisoned_teacher=poisoned, poisoned_fp32=run(poisoned, student, torch.float32),
               stubbed=sorted(set(s_.split(".")[0] for s_ in finder.stubbed)))
    torch.save(out, os.path.join(OUT, "kd_loop.pt"))
    print("KD loop golden written: loss fp32", float(out["clean_fp32"]["loss"]), "bf16", float(out["clean_bf16"]["loss"]),
          "poisoned", float(out["poisoned_fp32"]["loss"]), out["poisoned_fp32"]["printed"], "hooks", n_hooks)


# ----------------------------------------------------------------------------- A2: the older projector variants (a12)
def golden_proj_legacy():
    """``model_internvl/proj.py`` imported UNMODIFIED: its absent third-party roots are stubbed, ``transformers`` (the
    real library: T5Stack / T5Config) is imported first so the stubs do not shadow it.  Proj, Proj2, Proj3 and the bare MLP / MLP2 /
    MLP_plus run on CPU in fp32 with synthetic weights written by state-dict key; inputs, weights and outputs are stored."""
    import contextlib
    import io
    import transformers  # noqa: F401
    from transformers import (AutoModel, AutoModelForCausalLM, AutoTokenizer, BertModel, BertTokenizer, CLIPTextModel,  # noqa: F401
                              CLIPTextModelWithProjection, CLIPTokenizer, MT5EncoderModel, T5Config, T5EncoderModel,
                              T5ForConditionalGeneration, T5Tokenizer, T5TokenizerFast)
    from transformers.models.t5.modeling_t5 import T5Stack  # noqa: F401
    finder = _StubFinder()
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, REF)
    try:
        ref = importlib.import_module("model_internvl.proj")
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split(".")[0] in finder.ROOTS + ("model_internvl",)]:
            sys.modules.pop(k)
    out = dict(stubbed=sorted(set(s_.split(".")[0] for s_ in finder.stubbed)), transformers_version=transformers.__version__)
    kw = dict(in_channels=3, input_dim=64, output_dim0=32, output_dim1=128, num_layers=2, num_heads=2, head_dim=64)
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 3, 32, 64, generator=g).bfloat16().float()
    for name in ("Proj", "Proj2", "Proj3"):
        with contextlib.redirect_stdout(io.StringIO()):  # the constructors print their T5Config
            m = getattr(ref, name)(**kw).eval()
        sd = synth_state(m, 42, std=0.08)
        sd = {k: (v * 0 + 1 + 0.1 * torch.randn(v.shape, generator=g) if ("norm" in k and k.endswith("weight")) else v) for k, v in sd.items()}
        sd = {k: v.bfloat16().float() for k, v in sd.items()}  # bf16-representable weights: the bf16 product loads them exactly
        m.load_state_dict(sd)
        with torch.no_grad():
            x1, x2 = m(x)
        out[name] = dict(kwargs=kw, state={k: v for k, v in sd.items() if "embed_tokens" not in k}, x=x, x1=x1, x2=x2)
    xs = torch.randn(2, 32, 64, generator=g).bfloat16().float()
    for name in ("MLP", "MLP2", "MLP_plus"):
        m = getattr(ref, name)(in_dim=64, out_dim=128, hidden_dim=128, out_dim1=32).eval()
        sd = {k: v.bfloat16().float() for k, v in synth_state(m, 43, std=0.08).items()}
        m.load_state_dict(sd)
        with torch.no_grad():
            x1, x2 = m(xs)
        out[name] = dict(state=sd, x=xs, x1=x1, x2=x2)
    torch.save(out, os.path.join(OUT, "proj_legacy.pt"))
    print("legacy projector goldens written (reference model_internvl/proj.py, transformers", transformers.__version__, "); stubbed:", out["stubbed"])


# ----------------------------------------------------------------------------- B2: LightControl trainer helpers
def golden_lightcontrol_helpers():
    """lightcontrol/train_lightcontrol.py imported with every absent third-party root stubbed; its module-level helpers
    (_prepare_latent_image_ids :383, _pack_latents :396, _unpack_latents :403, get_sigmas :412) are called on seeded inputs."""
    extra = ("transformers", "zhconv")
    finder = _StubFinder()
    finder.ROOTS = finder.ROOTS + extra
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("transformers", "lightcontrol", "utils")}
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, REF)
    try:
        ref = importlib.import_module("lightcontrol.train_lightcontrol")
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split(".")[0] in finder.ROOTS + ("lightcontrol", "utils")]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(2, 16, 8, 12, generator=g)
    packed = ref._pack_latents(lat, 2, 16, 8, 12)

    def make_sched(num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True):
        """FlowMatchEulerDiscreteScheduler.__init__ [D031, recalled]: timesteps = linspace(1, N, N)[::-1]; sigmas = timesteps / N; the
        static shift is applied ONLY when use_dynamic_shifting is False; .timesteps = sigmas * N.  The reference builds it with
        from_pretrained(FLUX.1-dev, subfolder="scheduler") (train_lightcontrol.py:495-499), whose scheduler_config.json [recalled:
        shift 3.0, use_dynamic_shifting true, base/max_shift 0.5/1.15] therefore yields the UNSHIFTED linear table."""
        class Sched:
            pass
        sch = Sched()
        ts = torch.from_numpy(__import__("numpy").linspace(1, num_train_timesteps, num_train_timesteps, dtype="float32")[::-1].copy())
        sig = ts / num_train_timesteps
        if not use_dynamic_shifting:
            sig = shift * sig / (1 + (shift - 1) * sig)
        sch.sigmas, sch.timesteps = sig, sig * num_train_timesteps
        return sch
    sch, sch_static = make_sched(), make_sched(use_dynamic_shifting=False)
    idx = torch.tensor([0, 17, 500, 999])
    out = dict(lat=lat, packed=packed, unpacked=ref._unpack_latents(packed, 64, 96, 16),
               ids=ref._prepare_latent_image_ids(2, 8, 12, "cpu", torch.float32), idx=idx,
               sigmas=ref.get_sigmas(sch.timesteps[idx], sch, "cpu", n_dim=4, dtype=torch.float32),
               sigmas_static_shift3=ref.get_sigmas(sch_static.timesteps[idx], sch_static, "cpu", n_dim=4, dtype=torch.float32),
               scheduler_config=dict(num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True),
               stubbed=sorted(set(s_.split(".")[0] for s_ in finder.stubbed)))
    torch.save(out, os.path.join(OUT, "lightcontrol_helpers.pt"))
    print("LightControl helper goldens written; stubbed third-party roots:", out["stubbed"])


# ----------------------------------------------------------------------------- C: block structure
def _fake_diffusers():
    def register_to_config(init):
        @functools.wraps(init)
        def inner(self, *args, **kwargs):
            sig = inspect.signature(init)
            names = [n for n in sig.parameters][1:]
            cfg = {n: sig.parameters[n].default for n in names}
            cfg.update(dict(zip(names, args)))
            cfg.update(kwargs)
            object.__setattr__(self, "config", SimpleNamespace(**cfg))
            init(self, *args, **kwargs)
        return inner

    class _Mixin:
        pass

    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__dict__.update(attrs)
        mods[name] = m
        return m

    logging = SimpleNamespace(get_logger=lambda name: SimpleNamespace(warning=print, info=print))
    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=_Mixin, register_to_config=register_to_config)
    mod("diffusers.loaders", FromOriginalModelMixin=type("F", (), {}), PeftAdapterMixin=type("P", (), {}))
    mod("diffusers.models")
    mod("diffusers.models.attention", FeedForward=fo.FeedForward)
    mod("diffusers.models.attention_processor", Attention=fo.Attention, AttentionProcessor=object,
        FluxAttnProcessor2_0=fo.FluxAttnProcessor2_0, FusedFluxAttnProcessor2_0=fo.FluxAttnProcessor2_0)
    mod("diffusers.models.modeling_utils", ModelMixin=nn.Module)
    mod("diffusers.models.normalization", AdaLayerNormContinuous=fo.AdaLayerNormContinuous,
        AdaLayerNormZero=fo.AdaLayerNormZero, AdaLayerNormZeroSingle=fo.AdaLayerNormZeroSingle)
    mod("diffusers.utils", USE_PEFT_BACKEND=False, is_torch_version=lambda *a: True, logging=logging,
        scale_lora_layers=lambda *a, **k: None, unscale_lora_layers=lambda *a, **k: None,
        BaseOutput=dict)
    mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda c: c)
    mod("diffusers.models.embeddings",
        CombinedTimestepGuidanceTextProjEmbeddings=fo.CombinedTimestepGuidanceTextProjEmbeddings,
        CombinedTimestepTextProjEmbeddings=fo.CombinedTimestepTextProjEmbeddings, FluxPosEmbed=fo.FluxPosEmbed,
        TimestepEmbedding=fo.TimestepEmbedding, Timesteps=fo.Timesteps)
    mod("diffusers.models.modeling_outputs", Transformer2DModelOutput=lambda sample: SimpleNamespace(sample=sample))
    mod("diffusers.models.resnet", Downsample2D=_Anything, ResnetBlock2D=_Anything)
    return mods


def golden_flux_structure():
    fakes = _fake_diffusers()
    sys.modules.update(fakes)
    sys.path.insert(0, os.path.join(REF, "lightcontrol"))
    try:
        ref_flux = importlib.import_module("lightcontrol_flux")  # the reference file, unmodified
    finally:
        sys.path.remove(os.path.join(REF, "lightcontrol"))
        for k in fakes:
            sys.modules.pop(k, None)
    out = {}
    for tag, guidance in (("schnell", False), ("dev", True)):
        cfg = dict(patch_size=1, in_channels=8, num_layers=2, num_single_layers=3, attention_head_dim=16,
                   num_attention_heads=2, joint_attention_dim=24, pooled_projection_dim=12,
                   guidance_embeds=guidance, axes_dims_rope=(4, 6, 6))
        ref = ref_flux.FluxTransformer2DModel(**cfg).eval()
        sd = synth_state(ref, 31 + guidance, std=0.2)
        ref.load_state_dict(sd)
        g = torch.Generator().manual_seed(32)
        B, hl, wl, S = 2, 4, 6, 5
        inp = dict(hidden_states=torch.randn(B, hl * wl, 8, generator=g),
                   encoder_hidden_states=torch.randn(B, S, 24, generator=g),
                   pooled_projections=torch.randn(B, 12, generator=g),
                   timestep=torch.tensor([1.0, 0.25]),
                   img_ids=fo.prepare_latent_image_ids(2 * hl, 2 * wl), txt_ids=torch.zeros(S, 3))
        if guidance:
            inp["guidance"] = torch.tensor([3.5, 1.0])
        hooks = [[], [], []]
        def two(m, i, o, hooks=hooks):  # same contract as train_qwenvl.py:186-214
            hooks[0].append(o[0])
            hooks[1].append(o[1])

        def one(m, i, o, hooks=hooks):
            hooks[2].append(o)

        for b in ref.transformer_blocks:
            b.attn.register_forward_hook(two)
        for b in ref.single_transformer_blocks:
            b.attn.register_forward_hook(one)
        with torch.no_grad():
            y = ref(**inp, control_nets=[], return_dict=False)  # bare tensor (lightcontrol_flux.py:549-550)
        out[tag] = dict(cfg=cfg, state=sd, inputs=inp, output=y,
                        hooks=[torch.stack(h, 1) for h in hooks])
    torch.save(out, os.path.join(OUT, "flux_structure.pt"))
    print("flux structure golden written")


# ----------------------------------------------------------------------------- C2: ControlNeXt (LightControl branch)
def golden_controlnext():
    """The reference's ControlNeXtModel class + its injection into the transformer, imported unmodified with the diffusers
    leaves (ResnetBlock2D, Downsample2D, Timesteps, TimestepEmbedding) taken from the oracle."""
    from oracle import controlnext_oracle as co
    fakes = _fake_diffusers()
    fakes["diffusers.models.resnet"].ResnetBlock2D = co.ResnetBlock2D
    fakes["diffusers.models.resnet"].Downsample2D = co.Downsample2D
    fakes["diffusers.utils"].BaseOutput = dict
    sys.modules.update(fakes)
    sys.modules.pop("lightcontrol_flux", None)
    sys.path.insert(0, os.path.join(REF, "lightcontrol"))
    try:
        ref_flux = importlib.import_module("lightcontrol_flux")  # the reference file, unmodified
    finally:
        sys.path.remove(os.path.join(REF, "lightcontrol"))
        for k in fakes:
            sys.modules.pop(k, None)
    g = torch.Generator().manual_seed(51)
    net = ref_flux.ControlNeXtModel().eval()
    sd = synth_state(net, 52, std=0.05)
    net.load_state_dict(sd)
    hint = torch.rand(2, 3, 64, 96, generator=g) * 2 - 1
    t = torch.tensor([1000.0, 250.0])
    with torch.no_grad():
        o = net(hint, t)
    # the 6.7 M-parameter state is regenerated from its seed by the tests (synth_state(net, 52, std=0.05)), not stored
    out = dict(seed=52, std=0.05, hint=hint, timestep=t, out=o["out"], scale=o["scale"], n_params=sum(v.numel() for v in sd.values()),
               keys=sorted(sd.keys()))
    # injection inside the reference transformer (2 double blocks, 2 control nets; hidden width 3072 is fixed by mid_convs[1])
    cfg = dict(patch_size=1, in_channels=8, num_layers=2, num_single_layers=0, attention_head_dim=128,
               num_attention_heads=24, joint_attention_dim=16, pooled_projection_dim=8, guidance_embeds=True,
               axes_dims_rope=(16, 56, 56))
    ref = ref_flux.FluxTransformer2DModel(**cfg).eval()
    tsd = synth_state(ref, 53, std=0.02)
    ref.load_state_dict(tsd)
    nets = nn.ModuleList([ref_flux.ControlNeXtModel().eval() for _ in range(1)])  # 1 net, 2 blocks: exercises i < len(control_nets)
    for i, n in enumerate(nets):
        n.load_state_dict(synth_state(n, 54 + i, std=0.05))
    hl, wl, S = 4, 6, 4  # 64 x 96 hint pixels -> 4 x 6 control tokens (one per 16 x 16 pixels) = the packed-latent grid
    inp = dict(hidden_states=torch.randn(2, hl * wl, 8, generator=g), encoder_hidden_states=torch.randn(2, S, 16, generator=g),
               pooled_projections=torch.randn(2, 8, generator=g), timestep=torch.tensor([1.0, 0.25]),
               img_ids=fo.prepare_latent_image_ids(2 * hl, 2 * wl), txt_ids=torch.zeros(S, 3), guidance=torch.tensor([3.5, 3.5]))
    with torch.no_grad():
        y = ref(**inp, guided_hint=hint, control_nets=nets, return_dict=False)
    out["transformer"] = dict(cfg=cfg, seeds=dict(transformer=53, nets=[54], std_t=0.02, std_n=0.05), inputs=inp,
                              output=y)
    torch.save(out, os.path.join(OUT, "controlnext.pt"))
    print("controlnext golden written:", tuple(o["out"].shape), tuple(y.shape))


# ----------------------------------------------------------------------------- D: torchtitan cross-check
def crosscheck_torchtitan():
    try:
        from torchtitan.experiments.flux.model import layers as tl
    except Exception as e:  # pragma: no cover
        json.dump({"available": False, "why": repr(e)}, open(os.path.join(OUT, "crosscheck.json"), "w"))
        return
    torch.manual_seed(0)
    D, H = 256, 2
    g = torch.Generator().manual_seed(41)
    ids = torch.cat([torch.zeros(8, 3), fo.prepare_latent_image_ids(8, 8)])
    pe = tl.EmbedND(dim=128, theta=10000, axes_dim=[16, 56, 56])(ids[None])
    rope = fo.rope_table(ids)
    vec = torch.randn(2, D, generator=g)
    img, txt = torch.randn(2, 16, D, generator=g), torch.randn(2, 8, D, generator=g)
    res = {"available": True}

    ob = fo.FluxTransformerBlock(D, H, 128)
    ob.load_state_dict(synth_state(ob, 42))
    tb = tl.DoubleStreamBlock(D, H, 4.0, qkv_bias=True)
    with torch.no_grad():
        a = ob.attn
        for pre, mod_, q, k, v, nq, nk, o, ff in (
                ("img", ob.norm1, a.to_q, a.to_k, a.to_v, a.norm_q, a.norm_k, a.to_out[0], ob.ff),
                ("txt", ob.norm1_context, a.add_q_proj, a.add_k_proj, a.add_v_proj, a.norm_added_q, a.norm_added_k,
                 a.to_add_out, ob.ff_context)):
            getattr(tb, pre + "_mod").lin.load_state_dict(mod_.linear.state_dict())
            at = getattr(tb, pre + "_attn")
            at.qkv.weight.copy_(torch.cat([q.weight, k.weight, v.weight]))
            at.qkv.bias.copy_(torch.cat([q.bias, k.bias, v.bias]))
            at.norm.query_norm.weight.copy_(nq.weight)
            at.norm.key_norm.weight.copy_(nk.weight)
            at.norm.query_norm.eps = at.norm.key_norm.eps = 1e-6
            at.proj.load_state_dict(o.state_dict())
            mlp = getattr(tb, pre + "_mlp")
            mlp[0].load_state_dict(ff.net[0].proj.state_dict())
            mlp[2].load_state_dict(ff.net[2].state_dict())
        c_o, x_o = ob(img, txt, vec, rope)
        x_t, c_t = tb(img, txt, vec, pe)
    res["double_img_maxabs"] = float((x_o - x_t).abs().max())
    res["double_txt_maxabs"] = float((c_o - c_t).abs().max())

    osb = fo.FluxSingleTransformerBlock(D, H, 128)
    osb.load_state_dict(synth_state(osb, 43))
    tsb = tl.SingleStreamBlock(D, H, 4.0)
    with torch.no_grad():
        a = osb.attn
        tsb.modulation.lin.load_state_dict(osb.norm.linear.state_dict())
        tsb.linear1.weight.copy_(torch.cat([a.to_q.weight, a.to_k.weight, a.to_v.weight, osb.proj_mlp.weight]))
        tsb.linear1.bias.copy_(torch.cat([a.to_q.bias, a.to_k.bias, a.to_v.bias, osb.proj_mlp.bias]))
        tsb.linear2.load_state_dict(osb.proj_out.state_dict())
        tsb.norm.query_norm.weight.copy_(a.norm_q.weight)
        tsb.norm.key_norm.weight.copy_(a.norm_k.weight)
        tsb.norm.query_norm.eps = tsb.norm.key_norm.eps = 1e-6
        h = torch.cat([txt, img], 1)
        res["single_maxabs"] = float((osb(h, vec, rope) - tsb(h, vec, pe)).abs().max())
        t = torch.tensor([1000.0, 250.0])
        res["timestep_sinusoid_maxabs"] = float(
            (fo.Timesteps(256, True, 0)(t) - tl.timestep_embedding(t / 1000, 256)).abs().max())
    json.dump(res, open(os.path.join(OUT, "crosscheck.json"), "w"), indent=1)
    print("torchtitan cross-check:", res)


# ----------------------------------------------------------------------------- E: MiniCPM resampler
def golden_resampler():
    import builtins
    import math
    import typing
    # names the reference file expects torch.nn.functional's star-import to leak
    builtins.List, builtins.DType, builtins.math = typing.List, int, math
    sys.path.insert(0, os.path.join(REF, "minicpm"))
    try:
        ref_rs = importlib.import_module("resampler")  # the reference file, unmodified
    finally:
        sys.path.remove(os.path.join(REF, "minicpm"))
    torch.manual_seed(0)
    m = ref_rs.Resampler(num_queries=8, embed_dim=256, num_heads=2, kv_dim=48, adaptive=True, max_size=(6, 7)).eval()
    sd = synth_state(m, 51, std=0.08)
    sd["query"] = torch.randn(8, 256, generator=torch.Generator().manual_seed(52)) * 0.5
    m.load_state_dict(sd)
    tgt = torch.tensor([[3, 4], [2, 7], [5, 1]])
    x = torch.randn(3, 14, 48, generator=torch.Generator().manual_seed(53))
    with torch.no_grad():
        y = m(x, tgt)
    torch.save(dict(state=sd, x=x, tgt_sizes=tgt, out=y, pos_embed=m.pos_embed.clone()), os.path.join(OUT, "resampler_small.pt"))
    print("resampler golden written", tuple(y.shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1:  # regenerate selected fixtures only:  python oracle/make_golden.py controlnext
        for name in sys.argv[1:]:
            globals()["golden_" + name]()
        sys.exit(0)
    golden_projector()
    golden_proj_legacy()
    golden_helpers()
    golden_kd_loop()
    golden_flux_structure()
    crosscheck_torchtitan()
    golden_resampler()
    golden_controlnext()
    golden_lightcontrol_helpers()
