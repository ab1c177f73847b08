"""GPU: mha32 (flash-style MHA, head_dim 32) and xattn1 (absorbed one-query cross-attention) through the C-ABI
against torch fp32 references.  bf16 inputs → tolerance 2e-2 abs on O(1) outputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("groups,Sq,Sk,masked", [(3, 118, 118, False), (2, 64, 64, False), (2, 32, 20, False),
                                                 (2, 412, 412, False), (3, 47, 47, True), (1, 256, 256, False),
                                                 (2, 3, 3, False), (1, 352, 352, True)])
def test_mha32(groups, Sq, Sk, masked):
    from vgqa_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(Sq * 13 + Sk)
    q = torch.randn(groups * Sq, 768, device="cuda", generator=g).bfloat16()
    kv = q if Sq == Sk else torch.randn(groups * Sk, 768, device="cuda", generator=g).bfloat16()
    O = torch.zeros(groups * Sq, 256, device="cuda", dtype=torch.bfloat16)
    mask = None
    if masked:
        mask = (torch.rand(groups, Sk, device="cuda", generator=g) < 0.3)
        mask[:, 0] = False
        mask_u8 = mask.to(torch.uint8).contiguous()
    scale = 1 / math.sqrt(32)
    _lib.check(L.vgqa_mha32(_lib.ptr(q), 768, _lib.ptr(kv[:, 256:]), 768, _lib.ptr(kv[:, 512:]), 768, _lib.ptr(O), 256,
                            groups, Sq, Sk, _lib.ptr(mask_u8) if masked else None, scale, _stream()))
    torch.cuda.synchronize()
    qf = q[:, :256].float().view(groups, Sq, 8, 32).transpose(1, 2)
    kf = kv[:, 256:512].float().view(groups, Sk, 8, 32).transpose(1, 2)
    vf = kv[:, 512:].float().view(groups, Sk, 8, 32).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2) * scale
    if masked:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (s.softmax(-1) @ vf).transpose(1, 2).reshape(groups * Sq, 256)
    err = (O.float() - ref).abs().max().item()
    assert err < 2e-2, err


@pytest.mark.parametrize("F,S,tok0,Mk,use_pos,use_kpos,use_mask,want_att",
                         [(70, 118, 49, 69, True, False, False, False),    # TimeDecoder-like
                          (70, 118, 0, 69, False, True, False, False),     # PosDecoder-like
                          (33, 118, 69, 49, False, False, False, True),    # SpatialActivation-like
                          (9, 412, 0, 216, False, True, False, False),
                          (9, 27, 12, 15, True, False, True, False),
                          (5, 352, 208, 144, False, False, False, True)])
def test_xattn1(F, S, tok0, Mk, use_pos, use_kpos, use_mask, want_att):
    from vgqa_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(F * 31 + Mk)
    mem_all = torch.randn(F, S, 256, device="cuda", generator=g).bfloat16()
    qt = (torch.randn(F, 8, 256, device="cuda", generator=g) / 4).bfloat16()
    pos = torch.randn(Mk, 256, device="cuda", generator=g).bfloat16() if use_pos else None
    q2 = torch.randn(F, 256, device="cuda", generator=g).bfloat16() if use_kpos else None
    kpos_all = torch.randn(Mk, 1536, device="cuda", generator=g).bfloat16() if use_kpos else None
    kpos = kpos_all[:, 512:768] if use_kpos else None
    mask = None
    if use_mask:
        mask = torch.rand(F, S, device="cuda", generator=g) < 0.3
        mask[:, 0] = False
        mask_u8 = mask.to(torch.uint8).contiguous()
    ctx = torch.zeros(F, 2048, device="cuda", dtype=torch.bfloat16)
    att = torch.zeros(F, Mk, device="cuda") if want_att else None
    scale = 0.125
    mem = mem_all[:, tok0:tok0 + Mk]
    _lib.check(L.vgqa_xattn1(_lib.ptr(qt), _lib.ptr(mem), S, F, Mk, _lib.ptr(pos), 0, _lib.ptr(q2), _lib.ptr(kpos), 1536,
                             0, _lib.ptr(mask_u8) if use_mask else None, S, scale, _lib.ptr(ctx), _lib.ptr(att), _stream()))
    torch.cuda.synchronize()
    memf = mem.float()
    keys = memf + (pos.float() if use_pos else 0)
    if use_pos:
        keys = keys.bfloat16().float()   # the kernel rounds mem+pos to bf16 before the MMA
    s = torch.einsum("fhc,fmc->fhm", qt.float(), keys)
    if use_kpos:
        s = s + torch.einsum("fhd,mhd->fhm", q2.float().view(F, 8, 32), kpos.float().reshape(Mk, 8, 32))
    s = s * scale
    if use_mask:
        s = s.masked_fill(mask[:, None, :Mk], float("-inf"))
    p = s.softmax(-1)
    ref = torch.einsum("fhm,fmc->fhc", p, memf).reshape(F, 2048)
    err = (ctx.float() - ref).abs().max().item()
    assert err < 3e-2, err
    if want_att:
        a = p.sum(1).sigmoid()
        a = (a - a.min(1, keepdim=True)[0]) / (a.max(1, keepdim=True)[0] - a.min(1, keepdim=True)[0] + 1e-6)
        assert (att - a).abs().max().item() < 2e-2


@pytest.mark.parametrize("F,S,tok0,Mk,use_bias,use_mask,want_att",
                         [(70, 118, 49, 69, True, False, False),      # decoders: positional terms as an additive table
                          (700, 118, 0, 69, True, False, False),      # more frames than workers: the per-warp ring wraps
                          (33, 118, 69, 49, False, False, True),      # SpatialActivation
                          (9, 27, 12, 15, True, True, False),
                          (300, 250, 20, 100, True, True, True),      # 65..128 memory tokens (8x8 .. 10x10 maps)
                          (5, 352, 208, 144, False, False, True),     # cfg-5 memory length
                          (3, 412, 0, 208, True, True, True),
                          (2, 9, 1, 3, True, False, True)])
def test_xattn_stream(F, S, tok0, Mk, use_bias, use_mask, want_att):
    """Warp-per-frame streaming kernel (TMA ring, register-resident softmax) vs torch fp32."""
    from vgqa_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(F * 17 + Mk)
    mem_all = torch.randn(F, S, 256, device="cuda", generator=g).bfloat16()
    qt = (torch.randn(F, 8, 256, device="cuda", generator=g) / 4).bfloat16()
    ldsb = (Mk + 7) // 8 * 8
    sb = (torch.randn(F, 8, ldsb, device="cuda", generator=g) * 2) if use_bias else None
    mask = None
    if use_mask:
        mask = torch.rand(F, S, device="cuda", generator=g) < 0.3
        mask[:, 0] = False
        mask_u8 = mask.to(torch.uint8).contiguous()
    ctx = torch.zeros(F, 2048, device="cuda", dtype=torch.bfloat16)
    att = torch.zeros(F, Mk, device="cuda") if want_att else None
    scale = 0.125
    mem = mem_all[:, tok0:tok0 + Mk]
    _lib.check(L.vgqa_xattn1_bias(_lib.ptr(qt), _lib.ptr(mem), S, F, Mk, _lib.ptr(sb), ldsb,
                                  _lib.ptr(mask_u8) if use_mask else None, S, scale, _lib.ptr(ctx), _lib.ptr(att), _stream()))
    torch.cuda.synchronize()
    memf = mem.float()
    s = torch.einsum("fhc,fmc->fhm", qt.float(), memf)
    if use_bias:
        s = s + sb[:, :, :Mk]
    s = s * scale
    if use_mask:
        s = s.masked_fill(mask[:, None, :Mk], float("-inf"))
    p = s.softmax(-1)
    ref = torch.einsum("fhm,fmc->fhc", p, memf).reshape(F, 2048)
    err = (ctx.float() - ref).abs().max().item()
    assert err < 3e-2, err
    if want_att:
        a = p.sum(1).sigmoid()
        a = (a - a.min(1, keepdim=True)[0]) / (a.max(1, keepdim=True)[0] - a.min(1, keepdim=True)[0] + 1e-6)
        assert (att - a).abs().max().item() < 2e-2


@pytest.mark.parametrize("F,S,masked", [(5, 118, False), (300, 118, False), (3, 128, False), (7, 47, True), (2, 3, False),
                                        (4, 27, False), (9, 118, True),
                                        # more than 128 tokens per frame: online-softmax variant (attn_tc_long.cu)
                                        (3, 129, False), (5, 352, False), (2, 412, True), (1, 256, False), (160, 200, True),
                                        (2, 700, False)])
def test_enc_attn_tcgen05(F, S, masked):
    """tcgen05/TMEM per-frame attention vs torch fp32, and vs the warp-MMA kernel on the same data."""
    from vgqa_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(F * 7 + S)
    qkv = torch.randn(F * S, 768, device="cuda", generator=g).bfloat16()
    mask_u8 = None
    if masked:
        mask = torch.rand(F, S, device="cuda", generator=g) < 0.3
        mask[:, 0] = False
        mask_u8 = mask.to(torch.uint8).contiguous()
    scale = 1 / math.sqrt(32)
    outs = []
    for use_tc in (1, 0):
        O = torch.full((F * S + 8, 256), 7.0, device="cuda", dtype=torch.bfloat16)   # canary rows after the end
        _lib.check(L.vgqa_enc_attn(_lib.ptr(qkv), _lib.ptr(O), F, S, _lib.ptr(mask_u8), scale, use_tc, _stream()))
        torch.cuda.synchronize()
        assert float((O[F * S:].float() - 7.0).abs().max()) == 0.0
        outs.append(O[:F * S].float())
    q = qkv[:, :256].float().view(F, S, 8, 32).transpose(1, 2)
    k = qkv[:, 256:512].float().view(F, S, 8, 32).transpose(1, 2)
    v = qkv[:, 512:].float().view(F, S, 8, 32).transpose(1, 2)
    s = q @ k.transpose(-1, -2) * scale
    if masked:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(F * S, 256)
    assert (outs[0] - ref).abs().max().item() < 2e-2
    assert (outs[1] - ref).abs().max().item() < 2e-2
