"""VAE decoder (reference call site infer/inference_qwenvl.py:75,:209-216).

CPU: the oracle restatement against the BFL-derived decoder that ships in this image (torchtitan), key-name surface of the
drop-in.  GPU: the new kernels against torch fp32, the drop-in ``AutoencoderKL.decode`` against the fp32 oracle on the same
weights (<= 1e-2 relative), and one full-size 1024 px decode.
"""
import pytest
import torch

from parity import check


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _remap_to_bfl(vo, dec_sd, bfl_sd, cfg):
    m = vo.bfl_key_map(cfg)
    out = {}
    for k, t in dec_sd.items():
        pref = max((p for p in m if k.startswith(p + ".")), key=len)
        bk = m[pref] + "." + k[len(pref) + 1:].replace("conv_shortcut", "nin_shortcut")
        out[bk] = t[:, :, None, None] if (bfl_sd[bk].dim() == 4 and t.dim() == 2) else t
    return out


def test_vae_oracle_encoder_matches_bfl_encoder():
    bfl = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from oracle import vae_oracle as vo
    torch.manual_seed(1)
    cfg = dict(vo.FLUX_VAE_CONFIG, block_out_channels=(32, 64, 64, 64))
    v = vo.AutoencoderKL(**cfg).eval()
    b = bfl.Encoder(resolution=64, in_channels=3, ch=32, ch_mult=[1, 2, 2, 2], num_res_blocks=2, z_channels=16).eval()
    m, bsd, new = vo.bfl_encoder_key_map(cfg), b.state_dict(), {}
    for k, t in v.encoder.state_dict().items():
        pref = max((p for p in m if k.startswith(p + ".")), key=len)
        bk = m[pref] + "." + k[len(pref) + 1:].replace("conv_shortcut", "nin_shortcut")
        new[bk] = t[:, :, None, None] if (bsd[bk].dim() == 4 and t.dim() == 2) else t
    b.load_state_dict(new, strict=True)
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        assert float((v.encoder(x) - b(x)).abs().max()) < 1e-4
        d = v.encode(x).latent_dist
        assert d.mode().shape == (2, 16, 8, 8) and torch.equal(d.mode(), v.encoder(x)[:, :16])


def test_vae_oracle_matches_bfl_decoder():
    """Independent sanity anchor of the unpinned diffusers leaf: same weights, BFL key layout, fp32."""
    bfl = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from oracle import vae_oracle as vo
    torch.manual_seed(0)
    cfg = dict(vo.FLUX_VAE_CONFIG, block_out_channels=(32, 64, 64, 64))
    v = vo.AutoencoderKLDecoder(**cfg).eval()
    b = bfl.Decoder(ch=32, out_ch=3, ch_mult=[1, 2, 2, 2], num_res_blocks=2, in_channels=3, resolution=64, z_channels=16).eval()
    b.load_state_dict(_remap_to_bfl(vo, v.decoder.state_dict(), b.state_dict(), cfg), strict=True)
    z = torch.randn(2, 16, 8, 8)
    with torch.no_grad():
        assert float((v.decode(z)[0] - b(z)).abs().max()) < 1e-4


def test_vae_dropin_surface_and_keys():
    from oracle import vae_oracle as vo
    from x2i_b200 import vae as xv
    from x2i_b200._lib import X2IError
    m = xv.AutoencoderKL()
    o = vo.AutoencoderKL()
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    assert "encoder.down_blocks.0.downsamplers.0.conv.weight" in m.state_dict() and "encoder.down_blocks.3.downsamplers.0.conv.weight" not in m.state_dict()
    assert m.state_dict()["encoder.conv_out.weight"].shape == (32, 512, 3, 3)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    assert "decoder.mid_block.attentions.0.to_out.0.weight" in m.state_dict()
    assert "decoder.up_blocks.2.resnets.0.conv_shortcut.weight" in m.state_dict()
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in m.state_dict()
    assert m.config.scaling_factor == 0.3611 and m.config.shift_factor == 0.1159
    assert 2 ** len(m.config.block_out_channels) == 16          # vae_scale_factor of infer/inference_qwenvl.py:209
    with pytest.raises(X2IError):                               # no CPU path
        m.decode(torch.zeros(1, 16, 8, 8))
    img = xv.VaeImageProcessor(vae_scale_factor=16).postprocess(torch.tensor([[[[-1.0, 1.0]], [[0.0, 3.0]], [[-3.0, 0.5]]]]), output_type="pt")
    assert torch.equal(img, torch.tensor([[[[0.0, 1.0]], [[0.5, 1.0]], [[0.0, 0.75]]]]))


gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import ops
    return ops


@gpu
@pytest.mark.parametrize("C,G,act", [(512, 32, 2), (256, 32, 2), (128, 32, 2), (128, 32, 0), (64, 2, 1)])
def test_groupnorm_channel_and_group_forms(ops, C, G, act):
    g = torch.Generator(device="cuda").manual_seed(C + G)
    x = (torch.randn(2, 24, 40, C, device="cuda", generator=g) * 2 + 0.5).bfloat16()
    ga = torch.randn(C, device="cuda", generator=g).bfloat16()
    be = torch.randn(C, device="cuda", generator=g).bfloat16()
    y = ops.groupnorm_nhwc(x, ga, be, G, 1e-6, act=act)
    ref = torch.nn.functional.group_norm(x.float().permute(0, 3, 1, 2), G, ga.float(), be.float(), 1e-6)
    ref = {0: ref, 1: torch.relu(ref), 2: torch.nn.functional.silu(ref)}[act].permute(0, 2, 3, 1)
    assert _rel(y, ref) < 4e-3
    assert torch.equal(y, ops.groupnorm_nhwc(x, ga, be, G, 1e-6, act=act))  # deterministic


@gpu
def test_upsample_softmax_and_fp32_gemm(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2, 5, 7, 64, device="cuda", generator=g).bfloat16()
    up = ops.upsample2x_nhwc(x)
    ref = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)                                        # pure copy: bit-exact
    for cols in (64, 1000, 16384):
        s = torch.randn(37, cols, device="cuda", generator=g) * 6
        p = ops.softmax_rows(s)
        assert _rel(p, torch.softmax(s, -1)) < 4e-3
        assert float((p.float().sum(-1) - 1).abs().max()) < 1e-2
    a = torch.randn(300, 512, device="cuda", generator=g).bfloat16()
    b = torch.randn(512, 512, device="cuda", generator=g).bfloat16()
    c = ops.linear_f32(a, b, alpha=0.125)
    ref = 0.125 * (a.float() @ b.float().t())
    assert c.dtype == torch.float32 and _rel(c, ref) < 1e-5                    # fp32 accumulate, fp32 store: no bf16 rounding
    b2 = torch.randn(96, 512, device="cuda", generator=g).bfloat16()            # single-CTA kernel path (N % 256 != 0)
    assert _rel(ops.linear_f32(a, b2), a.float() @ b2.float().t()) < 1e-5


def _pair(cfg, seed):
    from oracle import vae_oracle as vo
    from x2i_b200 import vae as xv
    torch.manual_seed(seed)
    o = vo.AutoencoderKL(**cfg).eval()
    with torch.no_grad():
        for n, p in o.named_parameters():  # non-trivial GroupNorm affine parameters (the default init is weight 1, bias 0)
            if "norm" in n:
                p.copy_(torch.randn_like(p) * 0.3 + (1.0 if n.endswith("weight") else 0.0))
        for p in o.parameters():           # the product stores bf16: give both sides the same representable weights
            p.copy_(p.bfloat16().float())
    m = xv.AutoencoderKL(**cfg)
    m.load_state_dict(o.state_dict())
    return o, m.to("cuda", torch.bfloat16).eval()


@gpu
@pytest.mark.parametrize("blocks,hw", [((128, 128, 256, 256), (8, 12)), ((64, 128, 256, 512), (4, 8))])
def test_vae_decode_matches_oracle(ops, blocks, hw):
    cfg = dict(block_out_channels=blocks, norm_num_groups=32 if blocks[0] >= 128 else 16)
    o, m = _pair(cfg, 3)
    g = torch.Generator().manual_seed(4)
    z = torch.randn(2, 16, *hw, generator=g).bfloat16()
    with torch.no_grad():
        ref = o.decode(z.float())[0]
        got = m.decode(z.cuda(), return_dict=False)[0]
    assert got.shape == ref.shape == (2, 3, hw[0] * 8, hw[1] * 8) and got.dtype == torch.bfloat16
    err = _rel(got, ref)
    o16 = o.to("cuda", torch.bfloat16)
    with torch.no_grad():
        eager = _rel(o16.decode(z.cuda())[0], ref)  # the reference's own bf16 path (torch eager / cuDNN) on the same weights
    print(f"vae decode rel err vs fp32 oracle: x2i_b200 {err:.4f}, eager bf16 {eager:.4f}")
    # ~45 bf16-rounded layers with random weights amplify rounding noise: the reference's own bf16 path deviates 2-3 % from
    # fp32 on these nets (measured: eager 0.030 / 0.030, x2i_b200 0.017 / 0.020).  The bar is BASELINE.md's 1e-2 or, where the
    # eager bf16 path itself misses it, no worse than that path.
    check(f"vae decode {tuple(hw)} latents", err, eager)
    assert torch.equal(got, m.decode(z.cuda())[0])


@gpu
def test_decode_latents_and_pipeline_tail_full_size(ops):
    """1024 px (BASELINE configs 3/5): packed latents [2, 4096, 64] -> images [2, 3, 1024, 1024] in [0, 1]; sample 1 of the
    batch equals the same sample decoded alone (the element-wise parity test is test_vae_decode_matches_oracle)."""
    from x2i_b200 import vae as xv
    from x2i_b200.flux import init_synthetic_
    m = xv.AutoencoderKL().to("cuda", torch.bfloat16).eval()
    init_synthetic_(m, seed=5, std=0.03)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.fill_(1.0)
    g = torch.Generator(device="cuda").manual_seed(6)
    lat = torch.randn(2, 4096, 64, device="cuda", generator=g).bfloat16()
    with torch.no_grad():
        img = xv.decode_latents(m, lat, 1024, 1024)
        one = xv.decode_latents(m, lat[1:], 1024, 1024)
    assert img.shape == (2, 3, 1024, 1024)
    assert torch.isfinite(img.float()).all() and float(img.min()) >= 0.0 and float(img.max()) <= 1.0
    assert _rel(img[1:], one) < 1e-2


@gpu
def test_pipeline_decodes_through_vae_like_the_reference_tail(ops):
    """FluxPipeline(vae=...)(output_type="pt") == the reference's manual tail (infer/inference_qwenvl.py:209-216) applied to the
    output_type="latent" result of the same seed; without a vae only "latent" is allowed."""
    from x2i_b200 import smoke, vae as xv
    from x2i_b200._lib import X2IError
    from x2i_b200.flux import init_synthetic_
    from x2i_b200.pipeline import FlowMatchEulerDiscreteScheduler, FluxPipeline
    cfg = smoke.tiny_config(False)
    model, _ = smoke.make_pair(cfg, seed=11)
    vae = xv.AutoencoderKL(block_out_channels=(128, 128, 256, 256)).to("cuda", torch.bfloat16).eval()
    init_synthetic_(vae, seed=12, std=0.05)
    g = torch.Generator(device="cuda").manual_seed(1)
    pe = torch.randn(1, 16, cfg["joint_attention_dim"], device="cuda", generator=g).bfloat16()
    po = torch.randn(1, cfg["pooled_projection_dim"], device="cuda", generator=g).bfloat16()
    kw = dict(prompt_embeds=pe, pooled_prompt_embeds=po, num_inference_steps=2, guidance_scale=3.5, height=64, width=96)
    with torch.no_grad():
        pipe = FluxPipeline(scheduler=FlowMatchEulerDiscreteScheduler(shift=1.0), transformer=model, vae=vae)
        img = pipe(**kw, output_type="pt", generator=torch.Generator(device="cuda").manual_seed(7)).images
        lat = pipe(**kw, output_type="latent", generator=torch.Generator(device="cuda").manual_seed(7)).images
        ref = xv.decode_latents(vae, lat, 64, 96)
        pil = pipe(**kw, output_type="pil", generator=torch.Generator(device="cuda").manual_seed(7)).images
        bare = FluxPipeline(scheduler=FlowMatchEulerDiscreteScheduler(shift=1.0), transformer=model)
        with pytest.raises(X2IError):
            bare(**kw, output_type="pt")
    assert img.shape == (1, 3, 64, 96) and float((img - ref.float()).abs().max()) < 5e-3  # fp32 vs bf16 denormalise of the same decode
    assert len(pil) == 1 and pil[0].size == (96, 64)


@gpu
@pytest.mark.parametrize("blocks,hw", [((128, 128, 256, 256), (64, 96)), ((64, 128, 256, 512), (32, 64))])
def test_vae_encode_matches_oracle(ops, blocks, hw):
    """vae.encode(x).latent_dist (train_lightcontrol.py:678): moments against the fp32 oracle, sample() = mean + std * noise."""
    cfg = dict(block_out_channels=blocks, norm_num_groups=32 if blocks[0] >= 128 else 16)
    o, m = _pair(cfg, 13)
    g = torch.Generator().manual_seed(14)
    x = (torch.rand(2, 3, *hw, generator=g) * 2 - 1).bfloat16()
    with torch.no_grad():
        ref = o.encoder(x.float())
        dist = m.encode(x.cuda()).latent_dist
        got = dist.parameters
        eager = _rel(o.to("cuda", torch.bfloat16).encoder(x.cuda()), ref)
    assert got.shape == ref.shape == (2, 32, hw[0] // 8, hw[1] // 8)
    err = _rel(got, ref)
    print(f"vae encode rel err vs fp32 oracle: x2i_b200 {err:.4f}, eager bf16 {eager:.4f}")
    check(f"vae encode {tuple(hw)} px", err, eager)
    assert torch.equal(dist.mode(), got[:, :16])
    s1 = dist.sample(generator=torch.Generator(device="cuda").manual_seed(3))
    s2 = dist.sample(generator=torch.Generator(device="cuda").manual_seed(3))
    assert torch.equal(s1, s2) and s1.shape == (2, 16, hw[0] // 8, hw[1] // 8)
    assert torch.isfinite(s1.float()).all()


def test_vae_from_pretrained_directory_layout(tmp_path):
    """AutoencoderKL.from_pretrained(path, subfolder="vae", torch_dtype=...) as infer/inference_qwenvl.py:75 calls it: config.json +
    diffusion_pytorch_model.{safetensors,bin}; keys outside encoder./decoder. (a checkpoint with quant convs) are ignored."""
    import json
    from oracle import vae_oracle as vo
    from x2i_b200 import vae as xv
    cfg = dict(block_out_channels=[64, 64, 128, 128], norm_num_groups=16, latent_channels=16, layers_per_block=2,
               scaling_factor=0.3611, shift_factor=0.1159, _class_name="AutoencoderKL", _diffusers_version="0.31.0")
    o = vo.AutoencoderKL(**{k: v for k, v in cfg.items() if not k.startswith("_")})
    sd = dict(o.state_dict())
    sd["quant_conv.weight"] = torch.zeros(32, 32, 1, 1)  # must be ignored
    d = tmp_path / "flux" / "vae"
    d.mkdir(parents=True)
    (d / "config.json").write_text(json.dumps(cfg))
    torch.save(sd, d / "diffusion_pytorch_model.bin")
    m = xv.AutoencoderKL.from_pretrained(str(tmp_path / "flux"), subfolder="vae", torch_dtype=torch.bfloat16)
    assert m.dtype == torch.bfloat16 and m.config.block_out_channels == (64, 64, 128, 128)
    for k, v in o.state_dict().items():
        assert torch.equal(m.state_dict()[k].float(), v.bfloat16().float()), k
    try:
        from safetensors.torch import save_file
    except ImportError:
        return
    (d / "diffusion_pytorch_model.bin").unlink()
    save_file({k: v.contiguous() for k, v in o.state_dict().items()}, str(d / "diffusion_pytorch_model.safetensors"))
    m2 = xv.AutoencoderKL.from_pretrained(str(tmp_path / "flux"), subfolder="vae")
    assert torch.equal(m2.state_dict()["decoder.conv_out.bias"], o.state_dict()["decoder.conv_out.bias"])
