"""ORACLE (test infrastructure only) for the MLLM prefill with all-layer hidden-state capture.

The arithmetic lives in a third-party dependency that is NOT under /root/reference: ``transformers`` (the reference pins
``transformers==4.49.0``, requirements.txt:12; this image ships 5.5.0), class ``Qwen2_5_VLTextModel`` behind
``Qwen2_5_VLForConditionalGeneration.generate(..., output_hidden_states=True)`` as called at
``/root/reference/infer/inference_qwenvl.py:176`` and ``/root/reference/train/train_qwenvl.py:773-775``.  The in-image library IS
the oracle: this file only fixes how it is driven so that it reproduces the reference's call:

  * text-only positions as ``Qwen2_5_VLModel.get_rope_index`` builds them (cumsum(attention_mask) - 1, padded tokens 1, the same
    value in all three M-RoPE sections);
  * ``hidden_states`` of the prefill step stacked on dim 1 -> [B, num_layers + 1, S, H] (``torch.stack(hs[0], dim=1)``,
    train_qwenvl.py:775; ``torch.cat(hs[0]).unsqueeze(0)`` for B = 1, inference_qwenvl.py:123);
  * attention backend ``sdpa`` (the default the reference gets): padded query rows produce a zero attention output there, the
    ``eager`` backend averages all values instead -- the product documents and follows the sdpa / flash behaviour.

PARITY UNPINNED by the reference itself: it holds no tests or golden vectors for this path.
"""
import torch


def build(config: dict, attn_implementation: str = "sdpa", seed: int = 0, empty_on: str = None):
    """Random-weight ``Qwen2_5_VLTextModel`` for `config` (keys of x2i_b200.mllm.QWEN2_5_VL_3B); parameter names equal the product's.
    empty_on="cuda": allocate the parameters uninitialised directly on that device (full-size models: the caller fills them)."""
    from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as m
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLTextConfig
    cfg = Qwen2_5_VLTextConfig(vocab_size=config["vocab_size"], hidden_size=config["hidden_size"],
                               intermediate_size=config["intermediate_size"], num_hidden_layers=config["num_hidden_layers"],
                               num_attention_heads=config["num_attention_heads"], num_key_value_heads=config["num_key_value_heads"],
                               rms_norm_eps=config["rms_norm_eps"], max_position_embeddings=32768,
                               rope_parameters={"rope_type": "default", "mrope_section": [16, 24, 24], "rope_theta": config["rope_theta"]})
    cfg._attn_implementation = attn_implementation
    torch.manual_seed(seed)
    if empty_on is None:
        return m.Qwen2_5_VLTextModel(cfg).eval()
    with torch.device("meta"):
        model = m.Qwen2_5_VLTextModel(cfg)
    model = model.to_empty(device=empty_on).eval()
    model.rotary_emb = type(model.rotary_emb)(config=cfg).to(empty_on)  # its inv_freq buffer is computed at construction
    return model


def text_positions(attention_mask):
    mask = attention_mask.to(torch.int64)
    pos = (mask.cumsum(-1) - 1).masked_fill(mask == 0, 1)
    return pos[None].expand(3, -1, -1)


@torch.no_grad()
def prefill_hidden_states(model, input_ids, attention_mask):
    out = model(input_ids=input_ids, attention_mask=attention_mask, position_ids=text_positions(attention_mask).to(input_ids.device),
                output_hidden_states=True, use_cache=False)
    return torch.stack(out.hidden_states, dim=1)


def build_qwen2(config: dict, attn_implementation: str = "sdpa", seed: int = 0):
    """Random-weight plain ``Qwen2Model`` (the language model inside InternVL2.5-4B = Qwen2.5-3B-Instruct and MiniCPM-o-2.6 = Qwen2.5-7B;
    the reference vendors the same class as ``model_internvl/modeling_qwen2.py:778``): parameter names equal the product's."""
    from transformers import Qwen2Config, Qwen2Model
    cfg = Qwen2Config(vocab_size=config["vocab_size"], hidden_size=config["hidden_size"], intermediate_size=config["intermediate_size"],
                      num_hidden_layers=config["num_hidden_layers"], num_attention_heads=config["num_attention_heads"],
                      num_key_value_heads=config["num_key_value_heads"], rms_norm_eps=config["rms_norm_eps"], max_position_embeddings=32768,
                      rope_theta=config["rope_theta"], tie_word_embeddings=False)
    cfg._attn_implementation = attn_implementation
    torch.manual_seed(seed)
    return Qwen2Model(cfg).eval()


@torch.no_grad()
def prefill_hidden_states_plain(model, input_ids, attention_mask, position_mode: str = "arange"):
    """The two ways the reference drives a plain Qwen2 language model: 'arange' = ``language_model(inputs_embeds=..., attention_mask=...,
    output_hidden_states=True)`` without position_ids (InternVL, modeling_internvl_chat.py:357-363); 'cumsum' = HF ``generate``'s
    ``prepare_inputs_for_generation`` rule (MiniCPM-o)."""
    pos = None
    if position_mode == "cumsum":
        mask = attention_mask.to(torch.int64)
        pos = (mask.cumsum(-1) - 1).masked_fill(mask == 0, 1)
    out = model(inputs_embeds=model.embed_tokens(input_ids), attention_mask=attention_mask, position_ids=pos, output_hidden_states=True,
                use_cache=False)
    return torch.stack(out.hidden_states, dim=1)
