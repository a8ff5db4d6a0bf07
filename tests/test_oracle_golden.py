"""CPU: the oracle restatements against the fixtures minted FROM THE REFERENCE by oracle/make_golden.py."""
import json
import os

import pytest
import torch

from oracle import flux_oracle as fo
from oracle import kd_oracle, proj_oracle


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_projector_small_all_branches(golden_dir):
    g = _load(golden_dir, "proj_small.pt")
    for tag, kw in (("cnn", dict(use_scale=False, use_cnn=True)), ("scale", dict(use_scale=True, use_cnn=False)),
                    ("mean", dict(use_scale=False, use_cnn=False))):
        m = proj_oracle.Proj7Exp(in_channels=5, kernel_size=5, input_dim=64, output_dim0=24, output_dim1=96, **kw)
        m.load_state_dict(g[tag]["state"])  # identical keys to the reference module
        with torch.no_grad():
            pooled, seq = m(g[tag]["x"])
        torch.testing.assert_close(pooled, g[tag]["pooled"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(seq, g[tag]["seq"], rtol=1e-5, atol=1e-6)


def test_projector_config1(golden_dir):
    """BASELINE.json configs[0]: projector fwd + MSE vs random T5 embeds, batch=1 seq=77, CPU."""
    from oracle.make_golden import synth_state
    g = _load(golden_dir, "proj_c1.pt")
    m = proj_oracle.create_proj("qwen3b", 37, use_scale=False, use_cnn=True)
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] == 28317342
    m.load_state_dict(synth_state(m, 21, std=0.02))
    x = torch.randn(1, 37, 77, 2048, generator=torch.Generator().manual_seed(0))
    t5 = torch.randn(1, 77, 4096, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        pooled, seq = m(x)
    torch.testing.assert_close(pooled, g["pooled"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(seq[:, ::7, ::13], g["seq_sub"], rtol=1e-4, atol=1e-5)
    assert abs(float(torch.nn.functional.mse_loss(seq, t5)) - g["mse"]) < 1e-5


def test_train_helpers(golden_dir):
    g = _load(golden_dir, "helpers.pt")
    torch.testing.assert_close(kd_oracle.normalize(g["norm_in"]), g["norm_out"], rtol=0, atol=0)
    assert torch.equal(fo.prepare_latent_image_ids(8, 12), g["ids_8x12"])
    assert torch.equal(fo.prepare_latent_image_ids(128, 128), g["ids_128"])
    assert torch.equal(fo.pack_latents(g["pack_in"]), g["pack_out"])
    assert fo.calculate_shift(4096) == g["shift_4096"]
    assert fo.calculate_shift(1024) == g["shift_1024"]
    assert fo.calculate_shift(4096, 256, 4096, 0.5, 1.15) == g["shift_dev_4096"]
    # unpack is the inverse of pack (train_lightcontrol.py:403-410)
    lat = g["pack_in"]
    assert torch.equal(fo.unpack_latents(fo.pack_latents(lat), 8 * 8, 12 * 8, 16), lat)


def test_flux_structure_matches_reference_classes(golden_dir):
    """Oracle transformer == the reference's in-tree block/transformer classes (same leaves)."""
    g = _load(golden_dir, "flux_structure.pt")
    for tag in ("schnell", "dev"):
        e = g[tag]
        m = fo.FluxTransformer2DModel(**e["cfg"]).eval()
        m.load_state_dict(e["state"])  # diffusers key names
        hooks = [[], [], []]

        def two(mod, i, o, hooks=hooks):
            hooks[0].append(o[0]); hooks[1].append(o[1])

        def one(mod, i, o, hooks=hooks):
            hooks[2].append(o)

        for b in m.transformer_blocks:
            b.attn.register_forward_hook(two)
        for b in m.single_transformer_blocks:
            b.attn.register_forward_hook(one)
        with torch.no_grad():
            y = m(**e["inputs"], return_dict=False)[0]
        torch.testing.assert_close(y, e["output"], rtol=1e-5, atol=1e-6)
        for mine, ref in zip(hooks, e["hooks"]):
            torch.testing.assert_close(torch.stack(mine, 1), ref, rtol=1e-5, atol=1e-6)


def test_torchtitan_crosscheck_recorded(golden_dir):
    r = json.load(open(os.path.join(golden_dir, "crosscheck.json")))
    if r.get("available"):
        assert max(r["double_img_maxabs"], r["double_txt_maxabs"], r["single_maxabs"]) < 1e-5
        assert r["timestep_sinusoid_maxabs"] == 0.0


def test_kd_loss_direction_and_guard():
    g = torch.Generator().manual_seed(3)
    t = [torch.randn(2, 5, 48, generator=g) for _ in range(3)]
    s = [torch.randn(2, 5, 48, generator=g) for _ in range(3)]
    loss, skipped = kd_oracle.kd_loss(t, s)
    assert skipped == [] and float(loss) > 0
    # KL(student || teacher): zero iff identical
    assert abs(float(kd_oracle.kd_loss(t, t)[0])) < 1e-6
    # explicit formula
    def term(a, b):
        pa = torch.softmax(kd_oracle.normalize(a) / 3, -1)
        pb = torch.softmax(kd_oracle.normalize(b) / 3, -1)
        return (pb * (pb.log() - pa.log())).sum() / a.shape[0]
    torch.testing.assert_close(loss, sum(term(a, b) for a, b in zip(t, s)), rtol=1e-5, atol=1e-6)
    # a layer producing nan is skipped, not propagated (train_qwenvl.py:606-609)
    s[1][0, 0, 0] = float("nan")
    loss2, skipped2 = kd_oracle.kd_loss(t, s)
    assert skipped2 == [1] and torch.isfinite(loss2)


def test_rope_identity_on_text_tokens():
    cos, sin = fo.rope_table(torch.zeros(4, 3))
    assert torch.equal(cos, torch.ones(4, 128)) and torch.equal(sin, torch.zeros(4, 128))
    ids = fo.prepare_latent_image_ids(8, 8)
    cos, sin = fo.rope_table(ids)
    assert torch.equal(cos[:, :16], torch.ones(16, 16))  # axis 0 is always 0


def test_resampler_oracle_matches_reference(golden_dir):
    from oracle import resampler_oracle as ro
    g = _load(golden_dir, "resampler_small.pt")
    m = ro.Resampler(num_queries=8, embed_dim=256, num_heads=2, kv_dim=48, adaptive=True, max_size=(6, 7)).eval()
    m.load_state_dict(g["state"])
    assert torch.equal(m.pos_embed, g["pos_embed"])  # index-derived table: bit-exact
    with torch.no_grad():
        y = m(g["x"], g["tgt_sizes"])
    torch.testing.assert_close(y, g["out"], rtol=1e-5, atol=1e-6)


def test_adaln_leaves_match_an_independent_in_image_copy():
    """Second anchor for the un-vendored diffusers AdaLN leaves (parity is unpinned by the reference itself): the
    F5-TTS-derived DiT inside transformers' Qwen2.5-Omni carries the same two modules (chunk order shift/scale/gate x2, and the
    scale-FIRST order of the final AdaLayerNormContinuous).  Same weights -> identical outputs."""
    omni = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    if not hasattr(omni, "Qwen2_5_OmniAdaLayerNormZero"):
        pytest.skip("this transformers build has no Qwen2_5_OmniAdaLayerNormZero")
    from oracle import flux_oracle as fo
    torch.manual_seed(0)
    dim = 64
    mine, theirs = fo.AdaLayerNormZero(dim).eval(), omni.Qwen2_5_OmniAdaLayerNormZero(dim).eval()
    theirs.load_state_dict({k: v for k, v in mine.state_dict().items() if k.startswith("linear.")}, strict=False)
    x, emb = torch.randn(2, 7, dim), torch.randn(2, dim)
    with torch.no_grad():
        a, b = mine(x, emb), theirs(x, emb)
    assert len(a) == len(b) == 5
    for u, v in zip(a, b):
        assert torch.allclose(u, v, atol=1e-6)
    fin, fin_t = fo.AdaLayerNormContinuous(dim, dim).eval(), omni.Qwen2_5_OmniAdaLayerNormZero_Final(dim).eval()
    fin_t.load_state_dict({k: v for k, v in fin.state_dict().items() if k.startswith("linear.")}, strict=False)
    with torch.no_grad():
        assert torch.allclose(fin(x, emb), fin_t(x, emb), atol=1e-6)


def test_lightcontrol_trainer_helpers_match_reference(golden_dir):
    """The index / layout helpers of lightcontrol/train_lightcontrol.py (:383-:422), pinned by a fixture minted from the reference
    file itself (oracle/make_golden.py::golden_lightcontrol_helpers): bit-exact pack / unpack / ids, and the trainer's sigma table
    equals what the reference's get_sigmas() looks up in the scheduler."""
    from x2i_b200 import train_lightcontrol as tl
    from x2i_b200.pipeline import FluxPipeline
    d = torch.load(os.path.join(golden_dir, "lightcontrol_helpers.pt"))
    assert torch.equal(FluxPipeline._pack_latents(d["lat"], 2, 16, 8, 12), d["packed"])
    assert torch.equal(FluxPipeline._unpack_latents(d["packed"], 64, 96, 16), d["unpacked"])
    assert torch.equal(d["unpacked"], d["lat"])                                        # round trip
    assert torch.equal(FluxPipeline._prepare_latent_image_ids(2, 8, 12, "cpu", torch.float32), d["ids"])
    assert torch.equal(fo.unpack_latents(d["packed"], 64, 96, 16), d["unpacked"]) and torch.equal(fo.pack_latents(d["lat"]), d["packed"])
    table = tl.train_sigmas(1000, 3.0)
    assert torch.allclose(table[d["idx"]], d["sigmas"].flatten(), rtol=0, atol=1e-7)
