"""-m "not gpu": a pure-PyTorch model of the LAGGED online soft-max the attention forward runs by default
(x2i_b200/csrc/attn_sm100.cuh::softmax_step_lagged) -- the algorithmic claims, checked without a GPU:

  * the running reference need not be the row max: with ANY reference that keeps p = 2^(s - m) finite, O / l equals softmax(s) V;
  * the reference of step j + 1 derived from step j's ROW SUM (m + log2(sum p), an upper bound of the tile max within log2(128) = 7)
    and applied one step late keeps p bounded by 2^(growth of one tile) -- so only a tile that outgrows every earlier key by more
    than 2^ATT_LAG_LIMIT can overflow, and exactly that case raises the redo flag;
  * P rounded to bf16 (what the tensor pipe consumes) costs the same accuracy as in the classic scheme.

The GPU kernel itself is compared with torch fp32 in tests/test_gpu_fullsize.py::test_attention_lagged_form_matches_fp32_and_default.
"""
import math

import torch

ATT_LAG_RESCALE = 16.0  # csrc/attn_sm100.cuh
ATT_LAG_LIMIT = 96.0
TILE = 128


def lagged_attention(q, k, v, p_bf16=True):
    """q [Lq, d], k / v [Lk, d] fp32 -> (out [Lq, d], redo flag, number of lagged rescales).  Mirrors the kernel step by step:
    classic first tile, then lagged tiles; exp2 domain with scale log2(e) / sqrt(d)."""
    sc = math.log2(math.e) / math.sqrt(q.shape[-1])
    s = (q @ k.T) * sc  # [Lq, Lk], log2 units
    n_tiles = (k.shape[0] + TILE - 1) // TILE
    m = s[:, :TILE].max(dim=1).values  # classic first step: the tile's true max
    m_next = torch.full_like(m, -math.inf)
    l = torch.zeros_like(m)
    o = torch.zeros(q.shape[0], v.shape[1])
    redo, rescales = False, 0
    for j in range(n_tiles):
        sj = s[:, j * TILE:(j + 1) * TILE]
        vj = v[j * TILE:(j + 1) * TILE]
        alpha = torch.ones_like(m)
        if j > 0:
            grow = (m_next - m) > ATT_LAG_RESCALE
            alpha = torch.where(grow, torch.exp2(m - m_next), alpha)
            m = torch.where(grow, m_next, m)
            rescales += int(grow.sum())
        p = torch.exp2(sj - m[:, None])
        step_sum = p.sum(dim=1)
        if p_bf16:
            p = p.bfloat16().float()
        o = o * alpha[:, None] + p @ vj
        l = l * alpha + step_sum
        if j > 0:  # lagged steps only: the proxy for the next step and the overflow guard
            lg = torch.log2(step_sum)
            m_next = m + lg
            redo = redo or bool((~(lg <= ATT_LAG_LIMIT)).any())
    return o / l[:, None], redo, rescales


def _ref(q, k, v):
    return torch.softmax((q @ k.T) / math.sqrt(q.shape[-1]), dim=-1) @ v


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_lagged_reference_gives_the_exact_softmax_on_random_and_growing_scores():
    g = torch.Generator().manual_seed(0)
    Lq, Lk, d = 64, 1100, 128  # ragged last tile
    q, k, v = (torch.randn(n, d, generator=g) for n in (Lq, Lk, Lk))
    out, redo, _ = lagged_attention(q, k, v, p_bf16=False)
    assert not redo and _rel(out, _ref(q, k, v)) < 1e-5
    out, redo, _ = lagged_attention(q, k, v)  # P in bf16, as on the tensor pipe
    assert not redo and _rel(out, _ref(q, k, v)) < 4e-3
    # scores that grow by ~2^42 per key tile: a lagged rescale at (almost) every step, no overflow
    ramp = (torch.arange(Lk, dtype=torch.float32) * 0.005)[:, None].expand(Lk, d)
    qg, kg = torch.full((Lq, d), 4.0) + 0.1 * q, ramp + 0.05 * k
    out, redo, rescales = lagged_attention(qg, kg, v, p_bf16=False)
    assert not redo and rescales >= Lq * 6 and _rel(out, _ref(qg, kg, v)) < 1e-4
    # late keys that dominate single rows by ~2^25
    k2 = k.clone()
    k2[600:616] = q[:16] * 1.5
    out, redo, _ = lagged_attention(q, k2, v, p_bf16=False)
    assert not redo and _rel(out, _ref(q, k2, v)) < 1e-5


def test_row_sum_proxy_bounds_the_tile_max_and_the_guard_catches_overflow():
    g = torch.Generator().manual_seed(1)
    q, k = torch.randn(32, 128, generator=g), torch.randn(512, 128, generator=g)
    s = (q @ k.T) * (math.log2(math.e) / math.sqrt(128))
    for j in range(4):
        sj = s[:, j * TILE:(j + 1) * TILE]
        proxy = torch.log2(torch.exp2(sj).sum(dim=1))  # reference 0
        mx = sj.max(dim=1).values
        assert bool((proxy >= mx - 1e-4).all()) and bool((proxy <= mx + 7.0 + 1e-4).all())
    # ~2^167 per tile: fp32 exponentials overflow to inf within one lagged step -> the guard must fire (the kernel then re-runs the CTA's items
    # with the classic step); a growth of ~2^84 per tile stays below the 2^96 limit and must NOT fire
    v = torch.randn(640, 128, generator=g)
    for slope, expect in ((0.02, True), (0.01, False)):
        ramp = (torch.arange(640, dtype=torch.float32) * slope)[:, None].expand(640, 128)
        out, redo, _ = lagged_attention(torch.full((8, 128), 4.0), ramp.contiguous(), v, p_bf16=False)
        assert redo == expect
        if not expect:
            assert _rel(out, _ref(torch.full((8, 128), 4.0), ramp.contiguous(), v)) < 1e-4
