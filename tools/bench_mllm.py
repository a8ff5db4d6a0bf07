"""MLLM prefill at BASELINE config 2's conditioning shape: Qwen2.5-VL-3B text decoder (random weights), one prompt
left-padded to 512 tokens -> text_embeddings [B, 37, 512, 2048] -> Proj7Exp.  Prints one JSON line (device-timed)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=1)
    a = ap.parse_args()
    from x2i_b200 import _lib, mllm, proj as xproj
    dev = torch.device("cuda")
    m = mllm.Qwen2_5_VLTextPrefill.synthetic(mllm.QWEN2_5_VL_3B, device=dev, seed=0)
    pm = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True).to(dev, torch.bfloat16)
    B, S = a.batch, 512
    g = torch.Generator(device=dev).manual_seed(1)
    ids = torch.randint(0, 151000, (B, S), device=dev, generator=g)
    mask = torch.ones(B, S, dtype=torch.long, device=dev)
    mask[:, :300] = 0  # a 212-token prompt, left-padded
    out = torch.empty(B, 37, S, 2048, device=dev, dtype=torch.bfloat16)

    def step():
        with torch.no_grad():
            te = m.prefill_hidden_states(ids, mask, out=out)
            return pm(te)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    cfg = mllm.QWEN2_5_VL_3B
    H, F, L, Hq, Hkv = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"], cfg["num_key_value_heads"]
    flops = B * S * L * (2 * H * (Hq + 2 * Hkv) * 128 + 2 * H * H + 6 * H * F) + B * L * 2 * S * S * H  # causal attention: half of 4 S^2 H
    print(json.dumps({"workload": "Qwen2.5-VL-3B text prefill with all-layer capture [B,37,512,2048] + Proj7Exp (infer/inference_qwenvl.py:176-179)",
                      "batch": B, "ms_per_prefill_plus_projector": ms, "approx_tflops": flops / ms / 1e9,
                      "gpu_launches_per_call": (_lib.launch_count() - n0) / a.steps}))


if __name__ == "__main__":
    main()
