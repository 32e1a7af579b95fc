"""Kernel-level GPU tests through the C ABI: attention, LayerNorm, deformable attention, top-k.
The checker is the CPU oracle's math (oracle/lwdetr_oracle.py) or plain fp32 torch of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DTYPES = [torch.float16, torch.bfloat16]


def _tol(dt):
    return 3e-3 if dt == torch.float16 else 2e-2


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("nseq,seqlen,heads,dh", [(32, 100, 12, 16), (2, 1600, 12, 16), (16, 100, 12, 32), (1, 1600, 12, 32),
                                                  (16, 100, 12, 64), (1, 1600, 12, 64), (3, 300, 8, 32), (2, 100, 8, 32),
                                                  (2, 300, 12, 32), (5, 37, 4, 16),
                                                  # tcgen05 path (packed qkv, seqlen >= 512, dh >= 32): ragged tails, single-tile last CTA
                                                  (2, 700, 6, 32), (3, 640, 4, 64), (2, 520, 4, 32), (1, 1153, 4, 64)])
def test_attention_matches_softmax_reference(dt, nseq, seqlen, heads, dh):
    from b200 import capi
    g = torch.Generator(device="cuda").manual_seed(seqlen + dh)
    C = heads * dh
    qkv = (torch.randn(nseq * seqlen, 3 * C, device="cuda", generator=g) * 1.5).to(dt)
    out = torch.full((nseq * seqlen, C), float("nan"), device="cuda", dtype=dt)
    scale = dh ** -0.5
    capi.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, nseq, seqlen, heads, dh, scale)
    q, k, v = [t.float().reshape(nseq, seqlen, heads, dh).transpose(1, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
    ref = ((q * scale) @ k.transpose(-2, -1)).softmax(-1) @ v
    ref = ref.transpose(1, 2).reshape(nseq * seqlen, C)
    err = (out.float() - ref).abs().max().item()
    assert err <= _tol(dt) * ref.abs().max().item(), err


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("dh", [32, 64])
def test_attention_late_dominant_keys(dt, dh):
    """Keys whose scores dwarf everything seen before arrive late in the sequence: the lazily tracked row maximum of
    the tcgen05 kernel has to move (O rescaled in TMEM) several times, and rows must still normalise exactly."""
    from b200 import capi
    nseq, seqlen, heads = 2, 1600, 4
    g = torch.Generator(device="cuda").manual_seed(dh)
    C = heads * dh
    qkv = torch.randn(nseq * seqlen, 3 * C, device="cuda", generator=g)
    for start, gain in ((300, 4.0), (900, 10.0), (1500, 25.0)):
        for s0 in range(nseq):
            qkv[s0 * seqlen + start:s0 * seqlen + start + 7, C:2 * C] *= gain
    qkv = qkv.to(dt)
    out = torch.full((nseq * seqlen, C), float("nan"), device="cuda", dtype=dt)
    scale = dh ** -0.5
    capi.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, nseq, seqlen, heads, dh, scale)
    q, k, v = [t.float().reshape(nseq, seqlen, heads, dh).transpose(1, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
    ref = ((q * scale) @ k.transpose(-2, -1)).softmax(-1) @ v
    ref = ref.transpose(1, 2).reshape(nseq * seqlen, C)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err <= _tol(dt) * ref.abs().max().item(), err


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("rows,C,eps", [(3200, 192, 1e-6), (1000, 256, 1e-5), (777, 384, 1e-6), (64, 768, 1e-6)])
def test_layernorm(dt, rows, C, eps):
    from b200 import capi
    g = torch.Generator(device="cuda").manual_seed(C)
    xb = (torch.randn(rows, C + 64, device="cuda", generator=g) * 2 + 0.5).to(dt)
    x = xb[:, :C]
    w = torch.randn(C, device="cuda", generator=g)
    b = torch.randn(C, device="cuda", generator=g)
    y = torch.empty(rows, C, device="cuda", dtype=dt)
    capi.layernorm(x, y, w, b, eps)
    ref = F.layer_norm(x.float(), (C,), w, b, eps)
    assert (y.float() - ref).abs().max().item() <= _tol(dt) * ref.abs().max().item()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("B,Lq,M,L,P,shapes", [(1, 2, 2, 2, 2, [(6, 4), (3, 2)]),      # models/ops/test.py:27-34 shapes
                                               (3, 300, 16, 1, 2, [(40, 40)]),
                                               (2, 300, 24, 2, 4, [(80, 80), (20, 20)]),
                                               (2, 100, 16, 1, 4, [(40, 40)])])
def test_msda_forward_matches_oracle(dt, B, Lq, M, L, P, shapes):
    from b200 import capi
    from oracle import lwdetr_oracle as orc
    g = torch.Generator().manual_seed(3 + Lq)
    D, d = 16, M * 16
    S = sum(h * w for h, w in shapes)
    nlayers = 3
    value_all = torch.randn(B, S, nlayers * d, generator=g).to(dt)          # 3 layers side by side, use the middle one
    offs = torch.randn(B * Lq, M * L * P * 2, generator=g) * 2.0
    logit = torch.randn(B * Lq, M * L * P, generator=g) * 2.0
    ol = torch.cat([offs, logit], 1).to(dt)
    ref_box = torch.rand(B * Lq, 4, generator=g) * torch.tensor([1.2, 1.2, 0.6, 0.6]) - torch.tensor([0.1, 0.1, 0.0, 0.0])
    out = torch.full((B * Lq, d), float("nan"), dtype=dt, device="cuda")
    vg = value_all.cuda()
    capi.msda_forward(vg.reshape(B * S, nlayers * d)[:, d:2 * d], ol.cuda(), ref_box.cuda(), out, B, S, Lq, M, L, P, shapes)
    # oracle on the same rounded inputs (ops/modules/ms_deform_attn.py:118-131 + msda core)
    value = value_all[..., d:2 * d].float().reshape(B, S, M, D)
    off = ol[:, :M * L * P * 2].float().reshape(B, Lq, M, L, P, 2)
    aw = ol[:, M * L * P * 2:].float().reshape(B, Lq, M, L * P).softmax(-1).reshape(B, Lq, M, L, P)
    rb = ref_box.reshape(B, Lq, 4)
    loc = rb[:, :, None, None, None, :2] + off / P * rb[:, :, None, None, None, 2:] * 0.5
    ref = orc.msda_core(value, shapes, loc, aw).reshape(B * Lq, d)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= _tol(dt) * max(ref.abs().max().item(), 1.0), err


@pytest.mark.parametrize("B,S,k", [(2, 1600, 300), (3, 6800, 300), (1, 1600, 100), (2, 50, 50)])
def test_topk_sorted_indices(B, S, k):
    from b200 import capi
    g = torch.Generator(device="cuda").manual_seed(S)
    score = torch.randn(B, S, device="cuda", generator=g)
    score[0, 5] = score[0, 9]                       # a tie: lower index first
    idx = capi.topk(score, k).long()
    vals, ref = torch.sort(score, dim=1, descending=True, stable=True)
    assert torch.equal(torch.gather(score, 1, idx), vals[:, :k])
    assert torch.equal(idx, ref[:, :k])
