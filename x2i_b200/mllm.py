"""MLLM prefill with all-layer hidden-state capture -- the producer of the alignment projector's input.

The reference obtains the projector input by running the MLLM once over the (padded) prompt and keeping the hidden state of EVERY
layer:

    output = qwen_encoder.generate(**inputs, max_new_tokens=.., output_hidden_states=True, return_dict_in_generate=True)
    text_embeddings = torch.cat(output["hidden_states"][0]).unsqueeze(0)          # infer/inference_qwenvl.py:176-179, :121-132
    text_embeddings = torch.stack(generated_ids["hidden_states"][0], dim=1)       # train/train_qwenvl.py:773-775

``hidden_states[0]`` is the prefill step: (embeddings, output of layer 0, ..., output of layer L-2, final-norm(output of layer L-1)) --
``num_hidden_layers + 1`` tensors [B, S, H] (37 x 2048 for Qwen2.5-VL-3B, 29 x 3584 for the 7B model).  The model code itself is the
third-party ``transformers`` package (``Qwen2_5_VLTextModel``); this file is its text-only prefill on the x2i_b200 kernels:

  * every layer's output is written by the residual GEMM epilogue STRAIGHT INTO its slot of one ``[B, C, S, H]`` buffer -- the layout
    ``Proj7Exp.forward`` consumes -- so the reference's ``torch.cat`` / ``torch.stack`` of 78-106 MB never happens;
  * per layer: RMSNorm -> fused QKV GEMM (+bias) -> rotate-half RoPE + head-major split -> causal grouped-query attention with
    left-padding (the fused tcgen05 attention kernel, tiles right of the diagonal skipped) -> o_proj + residual -> RMSNorm ->
    gate/up GEMM with the SwiGLU epilogue -> down_proj + residual.

The language models of the reference's other two MLLM families are the same decoder: InternVL2.5-4B wraps Qwen2.5-3B-Instruct
(``infer/inference_internvl.py:76,:178-184``: its patched ``generate`` returns ``self.language_model(inputs_embeds, attention_mask,
output_hidden_states=True).hidden_states`` -- 37 x 2048, positions = ``arange(S)`` because no ``position_ids`` are passed,
``model_internvl/internvl/modeling_internvl_chat.py:357-363``), MiniCPM-o-2.6 wraps Qwen2.5-7B (``infer/inference_minicpm.py:78,:174-176``:
HF ``generate`` -> 29 x 3584, positions = cumsum(mask) - 1).  ``INTERNVL2_5_4B_LLM`` / ``MINICPM_O_2_6_LLM`` below are those configurations;
``position_mode`` selects the position rule (RoPE is relative, so both rules give the same states for the real tokens up to rounding).

Scope: text-only prompts (BASELINE config 2).  Image / video inputs need the vision tower, which stays with ``transformers``.
Padded positions: query rows with no visible key get a zero attention output (what transformers' sdpa and flash paths produce; its
eager path averages all values instead -- the reference's result at padded positions depends on the attention backend it runs with).
There is no CPU fallback.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import X2IError

BF16 = torch.bfloat16


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise X2IError(f"{type(self).__name__} is a parameter holder inside the fused x2i_b200 prefill")


class RMSNorm(_Holder):
    def __init__(self, dim, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.variance_epsilon = eps


class Attention(_Holder):
    def __init__(self, hidden, heads, heads_kv):
        super().__init__()
        self.q_proj = nn.Linear(hidden, heads * 128, bias=True)
        self.k_proj = nn.Linear(hidden, heads_kv * 128, bias=True)
        self.v_proj = nn.Linear(hidden, heads_kv * 128, bias=True)
        self.o_proj = nn.Linear(heads * 128, hidden, bias=False)


class MLP(_Holder):
    def __init__(self, hidden, inter):
        super().__init__()
        self.gate_proj = nn.Linear(hidden, inter, bias=False)
        self.up_proj = nn.Linear(hidden, inter, bias=False)
        self.down_proj = nn.Linear(inter, hidden, bias=False)


class DecoderLayer(_Holder):
    def __init__(self, hidden, inter, heads, heads_kv, eps):
        super().__init__()
        self.self_attn = Attention(hidden, heads, heads_kv)
        self.mlp = MLP(hidden, inter)
        self.input_layernorm = RMSNorm(hidden, eps)
        self.post_attention_layernorm = RMSNorm(hidden, eps)


QWEN2_5_VL_3B = dict(vocab_size=151936, hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16,
                     num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=1000000.0)
QWEN2_5_VL_7B = dict(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28,
                     num_key_value_heads=4, rms_norm_eps=1e-6, rope_theta=1000000.0)
# The language models inside the reference's other MLLMs (vocabulary sizes as published in the model repositories' config.json, recalled:
# no network here -- a checkpoint's own config wins).  InternVL2.5-1B's Qwen2.5-0.5B has head_dim 64 and is outside the d = 128 kernel.
INTERNVL2_5_4B_LLM = dict(vocab_size=151674, hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16,
                          num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=1000000.0, position_mode="arange")
MINICPM_O_2_6_LLM = dict(vocab_size=151700, hidden_size=3584, intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28,
                         num_key_value_heads=4, rms_norm_eps=1e-6, rope_theta=1000000.0, position_mode="cumsum")


class Qwen2_5_VLTextPrefill(nn.Module):
    """Parameter names follow ``transformers``' ``Qwen2_5_VLTextModel`` (``embed_tokens``, ``layers.N.self_attn.q_proj`` ...,
    ``norm``); ``load_hf_state_dict`` accepts a ``Qwen2_5_VLForConditionalGeneration`` checkpoint (either key generation)."""

    def __init__(self, vocab_size=151936, hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16,
                 num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=1000000.0, head_dim=128, position_mode="cumsum"):
        super().__init__()
        if position_mode not in ("cumsum", "arange"):
            raise X2IError("Qwen2_5_VLTextPrefill: position_mode is 'cumsum' (Qwen2.5-VL get_rope_index / HF generate) or 'arange' "
                           "(a plain forward without position_ids: InternVL's patched generate)")
        self.position_mode = position_mode
        if head_dim != 128 or hidden_size // num_attention_heads != 128:
            raise X2IError("Qwen2_5_VLTextPrefill: the attention kernel is specialised for head_dim 128 (all Qwen2.5-VL sizes)")
        if intermediate_size % 128 or hidden_size % 64 or num_attention_heads % num_key_value_heads:
            raise X2IError("Qwen2_5_VLTextPrefill: intermediate_size % 128, hidden_size % 64 and heads % kv_heads must be 0")
        self.config = SimpleNamespace(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                                      num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                                      num_key_value_heads=num_key_value_heads, rms_norm_eps=rms_norm_eps, rope_theta=rope_theta)
        self.embed_tokens = nn.Embedding(vocab_size, hidden_size)
        self.layers = nn.ModuleList([DecoderLayer(hidden_size, intermediate_size, num_attention_heads, num_key_value_heads, rms_norm_eps)
                                     for _ in range(num_hidden_layers)])
        self.norm = RMSNorm(hidden_size, rms_norm_eps)
        self._packed = None

    # ---- weights ---------------------------------------------------------------------------------------------------------
    @classmethod
    def synthetic(cls, config: dict, device="cuda", seed: int = 0, std: float = 0.02):
        from .flux import init_synthetic_
        with torch.device("meta"):
            m = cls(**config)
        m = m.to(BF16).to_empty(device=device)
        return init_synthetic_(m, seed=seed, std=std).eval()

    def load_hf_state_dict(self, state_dict, strict: bool = True):
        """Load the language-model part of a Qwen2.5-VL checkpoint: keys ``model.language_model.*`` (transformers >= 4.52),
        ``model.*`` (4.49, the reference's pin) or bare; ``visual.*`` and ``lm_head.*`` are ignored."""
        sd = {}
        for k, v in state_dict.items():
            for pre in ("model.language_model.", "language_model.model.", "language_model.", "llm.model.", "llm.", "model."):
                if k.startswith(pre):
                    k = k[len(pre):]
                    break
            if k.startswith(("visual.", "lm_head.", "model.visual.", "vision_model.", "mlp1.", "vpm.", "resampler.", "apm.", "tts.",
                             "audio_projection_layer.", "audio_avg_pooler.")):
                continue
            sd[k] = v
        self._packed = None
        return self.load_state_dict(sd, strict=strict)

    def _pack(self):
        w0 = self.layers[0].self_attn.q_proj.weight
        if self._packed is not None and self._packed[0] == (w0.data_ptr(), w0._version):
            return self._packed[1]
        if w0.dtype != BF16 or not w0.is_cuda:
            raise X2IError("Qwen2_5_VLTextPrefill runs in bf16 on a CUDA device: call .to('cuda', torch.bfloat16) first")
        packs = []
        for l in self.layers:
            a, m = l.self_attn, l.mlp
            packs.append(dict(w_qkv=torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).contiguous(),
                              b_qkv=torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).contiguous(),
                              w_gu=ops.pack_swiglu_weight(m.gate_proj.weight.detach(), m.up_proj.weight.detach())))
        # Qwen2_5_VLRotaryEmbedding (default rope): inv_freq = 1 / theta^(2i / 128), fp32 (index work: computed once, exactly as there)
        inv_freq = 1.0 / (self.config.rope_theta ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))
        self._packed = ((w0.data_ptr(), w0._version), dict(layers=packs, inv_freq=inv_freq.to(w0.device)))
        return self._packed[1]

    # ---- prefill ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def text_positions(attention_mask, mode: str = "cumsum"):
        """mode 'cumsum': Qwen2_5_VLModel.get_rope_index for text-only inputs (and HF generate's rule for plain Qwen2): position =
        cumsum(mask) - 1, padded tokens 1 (all three M-RoPE sections carry it).  mode 'arange': position = column index, what a forward
        without position_ids uses (InternVL's patched generate).  Also the first valid key per row for the attention kernel (left padding)."""
        mask = attention_mask.to(torch.int64)
        if mode == "arange":
            pos = torch.arange(mask.shape[1], device=mask.device, dtype=torch.int64)[None].expand_as(mask)
        else:
            pos = (mask.cumsum(-1) - 1).masked_fill(mask == 0, 1)
        if bool(((mask[:, 1:] - mask[:, :-1]) < 0).any()):
            raise X2IError("Qwen2_5_VLTextPrefill: only left padding is supported (the reference pads on the left, "
                           "train/train_qwenvl.py:397); a 1 -> 0 transition was found in attention_mask")
        start = (mask.shape[1] - mask.sum(-1)).to(torch.int32)
        return pos.to(torch.int32).contiguous(), start.contiguous()

    @torch.no_grad()
    def prefill_hidden_states(self, input_ids, attention_mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """input_ids [B, S] int64, attention_mask [B, S] (1 = token, left-padded) -> text_embeddings [B, C, S, H] bf16 with
        C = num_hidden_layers + 1, exactly ``torch.stack(generate(...).hidden_states[0], dim=1)``."""
        cfg = self.config
        self._pack()
        dev = self.embed_tokens.weight.device
        input_ids = input_ids.to(dev)
        B, S = input_ids.shape
        H, L = cfg.hidden_size, cfg.num_hidden_layers
        if attention_mask is None:
            attention_mask = torch.ones(B, S, dtype=torch.int64, device=dev)
        pos, start = self.text_positions(attention_mask.to(dev), self.position_mode)
        if out is not None and (out.shape != (B, L + 1, S, H) or out.dtype != BF16 or not out.is_contiguous()):
            raise X2IError("prefill_hidden_states: out must be a contiguous bf16 [B, num_layers + 1, S, hidden] tensor")
        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            if out is None:
                out = torch.empty(B, L + 1, S, H, device=dev, dtype=BF16)
            return self._prefill_body(input_ids, pos, start, out)
        # One captured graph per (B, S): ~8 launches per layer are host-bound when issued one by one (295 launches in 6.4 ms for the 3B
        # model against ~3.5 ms of GPU time).  Inputs are copied into the graph's static buffers, the capture buffer is copied out.
        w0 = self.layers[0].self_attn.q_proj.weight
        key = (B, S, w0.data_ptr(), w0._version)
        st = self._graphs.get(key)
        if st is None:
            sin = dict(ids=input_ids.clone(), pos=pos.clone(), start=start.clone(),
                       out=torch.empty(B, L + 1, S, H, device=dev, dtype=BF16))
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up outside the capture: function attributes, workspaces
                self._prefill_body(sin["ids"], sin["pos"], sin["start"], sin["out"])
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                self._prefill_body(sin["ids"], sin["pos"], sin["start"], sin["out"])
            st = (graph, sin, _lib.launch_count() - n0)
            self._graphs = {key: st}  # one shape resident
        graph, sin, n_kernels = st
        sin["ids"].copy_(input_ids, non_blocking=True)
        sin["pos"].copy_(pos, non_blocking=True)
        sin["start"].copy_(start, non_blocking=True)
        graph.replay()
        _lib.note_graph_replay(n_kernels)
        if out is None:
            return sin["out"].clone()
        out.copy_(sin["out"])
        return out

    def _prefill_body(self, input_ids, pos, start, out):
        cfg = self.config
        pk = self._pack()
        dev = out.device
        B, S = input_ids.shape
        H, Hq, Hkv, L = cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.num_hidden_layers
        ops.gather_rows(input_ids, self.embed_tokens.weight, out[:, 0])
        xn = torch.empty(B, S, H, device=dev, dtype=BF16)
        qkv = torch.empty(B, S, (Hq + 2 * Hkv) * 128, device=dev, dtype=BF16)
        q = torch.empty(B, Hq, S, 128, device=dev, dtype=BF16)
        k = torch.empty(B, Hkv, S, 128, device=dev, dtype=BF16)
        v = torch.empty_like(k)
        att = torch.empty(B, S, Hq * 128, device=dev, dtype=BF16)
        mid = torch.empty(B, S, H, device=dev, dtype=BF16)
        act = torch.empty(B, S, cfg.intermediate_size, device=dev, dtype=BF16)
        last = torch.empty(B, S, H, device=dev, dtype=BF16)
        for i, layer in enumerate(self.layers):
            w = pk["layers"][i]
            h_in = out[:, i]                                     # [B, S, H] view: this layer's input lives in its capture slot
            h_out = out[:, i + 1] if i + 1 < L else last         # the last layer's output is captured after the final norm
            ops.rmsnorm(h_in, layer.input_layernorm.weight, cfg.rms_norm_eps, out=xn)
            ops.linear(xn, w["w_qkv"], w["b_qkv"], out=qkv)
            ops.rope_half_split(qkv, pos, pk["inv_freq"], Hq, Hkv, q=q, k=k, v=v)
            ops.causal_attention(q, k, v, kv_start=start, out=att)
            for b in range(B):                                   # residual rows of a capture slot are batch-strided: one GEMM per sample
                ops.linear_residual(att[b], layer.self_attn.o_proj.weight, h_in[b], mid[b])
            ops.rmsnorm(mid, layer.post_attention_layernorm.weight, cfg.rms_norm_eps, out=xn)
            ops.linear_swiglu(xn, w["w_gu"], out=act)
            for b in range(B):
                ops.linear_residual(act[b], layer.mlp.down_proj.weight, mid[b], h_out[b])
        ops.rmsnorm(last, self.norm.weight, cfg.rms_norm_eps, out=out[:, L])
        return out

    use_cuda_graph = True  # replay one captured graph per (B, S) instead of ~8 launches per layer
    _graphs = {}

    def generate(self, input_ids=None, attention_mask=None, max_new_tokens: int = 1, output_hidden_states: bool = True,
                 return_dict_in_generate: bool = True, **unused):
        """The call shape of the reference (train/train_qwenvl.py:773-774, infer/inference_qwenvl.py:176): returns an object whose
        ``["hidden_states"][0]`` is the tuple of per-layer prefill hidden states (views of one [B, C, S, H] buffer, also exposed as
        ``.text_embeddings`` so callers can skip the stack).  Token generation itself is out of scope: ``sequences`` is the prompt."""
        if not (output_hidden_states and return_dict_in_generate):
            raise X2IError("Qwen2_5_VLTextPrefill.generate: only the hidden-state capture form of the reference is provided "
                           "(output_hidden_states=True, return_dict_in_generate=True)")
        te = self.prefill_hidden_states(input_ids, attention_mask)
        hs = (tuple(te[:, c] for c in range(te.shape[1])),)
        return _GenerateOutput(sequences=input_ids, hidden_states=hs, text_embeddings=te)

    def forward(self, input_ids, attention_mask=None):
        return self.prefill_hidden_states(input_ids, attention_mask)


class _GenerateOutput(dict):
    """dict- and attribute-style access like transformers' ModelOutput."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.__dict__.update(kw)
