"""the older projector variants of model_internvl/proj.py (MLP, MLP2, MLP_plus, Proj, Proj2, Proj3) against a
fixture minted by importing the reference file itself (oracle/make_golden.py::golden_proj_legacy; the T5Stack inside is the real
``transformers`` one).  CPU: naming / index work; -m gpu: the drop-ins on the x2i_b200 kernels."""
import os

import pytest
import torch

from parity import record

gpu = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


def test_state_dict_names_match_the_reference_classes(golden_dir):
    from x2i_b200 import proj_legacy as pl
    d = torch.load(os.path.join(golden_dir, "proj_legacy.pt"))
    for name in ("Proj", "Proj2", "Proj3"):
        m = getattr(pl, name)(**d[name]["kwargs"])
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in d[name]["state"].items()}, name
        sd = dict(d[name]["state"])
        sd["t5stack.embed_tokens.weight"] = torch.zeros(4, 4)  # the library's unused table in a real checkpoint: accepted, dropped
        m.load_state_dict(sd)
    for name in ("MLP", "MLP2", "MLP_plus"):
        m = getattr(pl, name)(in_dim=64, out_dim=128, hidden_dim=128, out_dim1=32)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in d[name]["state"].items()}, name


def test_relative_position_bucket_matches_transformers():
    from transformers.models.t5.modeling_t5 import T5Attention
    from x2i_b200.proj_legacy import relative_position_bucket
    ctx, mem = torch.arange(300)[:, None], torch.arange(300)[None, :]
    want = T5Attention._relative_position_bucket(mem - ctx, bidirectional=True, num_buckets=32, max_distance=128)
    assert torch.equal(relative_position_bucket(mem - ctx, 32, 128), want)


@gpu
@pytest.mark.parametrize("name", ["MLP", "MLP2", "MLP_plus", "Proj", "Proj2", "Proj3"])
def test_legacy_projectors_match_the_reference_file(golden_dir, name):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import proj_legacy as pl
    d = torch.load(os.path.join(golden_dir, "proj_legacy.pt"))[name]
    m = getattr(pl, name)(**d["kwargs"]) if "kwargs" in d else getattr(pl, name)(in_dim=64, out_dim=128, hidden_dim=128, out_dim1=32)
    m.load_state_dict(d["state"])
    m = m.to("cuda", torch.bfloat16).eval()
    with torch.no_grad():
        x1, x2 = m(d["x"].to("cuda", torch.bfloat16))
    assert x1.shape == d["x1"].shape and x2.shape == d["x2"].shape
    e1, e2 = rel(x1, d["x1"]), rel(x2, d["x2"])
    record(f"a12 {name} vs the reference file model_internvl/proj.py run in fp32 (same bf16-representable weights): pooled x1", e1)
    record(f"a12 {name} vs the reference file model_internvl/proj.py run in fp32 (same bf16-representable weights): sequence x2", e2)
    assert e2 < 1e-2 and e1 < 1e-2, (e1, e2)  # 1e-2 relative for bf16 paths
