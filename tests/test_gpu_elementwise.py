"""Memory-bound decoder glue through the C-ABI against the oracle's torch statement of the same ops."""
import pytest
import torch

from oracle import decoder as od
from tests.gpu_util import ctx, record, rel_err

pytestmark = pytest.mark.gpu
DT = torch.bfloat16


@pytest.mark.parametrize("t,h,s", [(1, 256, 0), (5, 5120, 0), (3, 1024, 4), (32, 5120, 7), (40, 264, 1)])
def test_reduce_residual_rmsnorm(t, h, s):
    _rmsnorm_case(t, h, s, DT)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("t,h", [(257, 256), (300, 5120), (1031, 1024), (513, 4096), (260, 6144)])
def test_plain_rmsnorm_of_many_rows(t, h, dt):
    """t > 256 rows and nothing to reduce (the prefill after a GEMM with the residual in its epilogue): one warp per row (rmsnorm_rows_kernel),
    rows not a multiple of the warps per CTA, widths that fill 1 .. 24 vectors per lane; also with resid_out == resid_in (the TP prefill)."""
    _rmsnorm_case(t, h, 0, dt)
    _rmsnorm_case(t, h, 0, dt, alias_out=True)


def _rmsnorm_case(t, h, s, DT, alias_out=False):
    c = ctx()
    g = torch.Generator().manual_seed(t * 7 + h + s)
    resid = (torch.randn(t, h, generator=g)).to(DT)
    w = (torch.rand(h, generator=g) + 0.5).to(DT)
    eps = 1e-6
    if s:
        part = torch.randn(s, t, h, generator=g) * 0.3
        proj = part.sum(0).to(DT)
        hh = (resid.float() + proj.float()).to(DT)
    else:
        part, hh = None, resid
    ref_norm = od.rms_norm(hh, w, eps)
    r_out = torch.empty(t, h, device="cuda", dtype=DT)
    n_out = torch.full((t, h), float("nan"), device="cuda", dtype=DT)
    rin = resid.cuda()
    c.reduce_residual_rmsnorm(part.cuda() if s else None, s, rin, (rin if alias_out else (r_out if s else None)), w.cuda(), eps, n_out)
    torch.cuda.synchronize()
    if s:
        assert rel_err(r_out, hh) < 8e-3
    e = rel_err(n_out, ref_norm)
    record("rmsnorm", t=t, h=h, s=s, err=e)
    assert e < 8e-3


def test_reduce_swiglu():
    c = ctx()
    t, inter, s = 6, 200, 3
    part = torch.randn(s, t, 2 * inter, generator=torch.Generator().manual_seed(8))
    out = torch.empty(t, inter, device="cuda", dtype=DT)
    c.reduce_swiglu(part.cuda(), s, t, inter, out, interleaved=False)
    tot = part.sum(0)
    g, u = tot[:, :inter].to(DT), tot[:, inter:].to(DT)
    ref = (torch.nn.functional.silu(g.float()).to(DT).float() * u.float()).to(DT)
    assert rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("partial", [True, False])
@pytest.mark.parametrize("d", [64, 128])
def test_qkv_rope_cache(partial, d):
    c = ctx()
    from chatts_b200.config import ChatTSConfig
    from chatts_b200.model import rope_tables
    t, nh, nkv, page = 11, 4, 2, 16
    cfg = ChatTSConfig.tiny(head_dim=d, num_attention_heads=nh, num_key_value_heads=nkv)
    width = (nh + 2 * nkv) * d
    g = torch.Generator().manual_seed(3)
    bias = (torch.randn(width, generator=g) * 0.1).to(DT)
    positions = torch.tensor([0, 1, 2, 5, 9, 100, 37, 3, 4, 63, 64], dtype=torch.int32)
    if partial:
        s = 3
        part = torch.randn(s, t, width, generator=g)
        qkv = (part.sum(0) + bias.float()).to(DT)
        src = part.cuda()
    else:
        s = 1
        qkv = torch.randn(t, width, generator=g).to(DT)
        src = qkv.cuda()
    cos, sin = rope_tables(cfg, 128, DT, "cuda")
    n_pages = 8
    kc = torch.zeros(n_pages, nkv, page, d, device="cuda", dtype=DT)
    vc = torch.zeros_like(kc)
    slot = torch.tensor([5, 6, 7, 40, 41, 100, 17, 18, -1, 127, 0], dtype=torch.int32)
    q_out = torch.empty(t, nh * d, device="cuda", dtype=DT)
    k_lin = torch.empty(t, nkv * d, device="cuda", dtype=DT)
    v_lin = torch.empty(t, nkv * d, device="cuda", dtype=DT)
    c.qkv_rope_cache(src, partial, s, bias.cuda() if partial else None, positions.cuda(), cos, sin, slot.cuda(), q_out, kc, vc,
                     k_lin, v_lin, t, nh, nkv, d, page)
    torch.cuda.synchronize()
    # oracle: HF rope with full [T, d] tables
    oc, osn = od.rope_tables(dict(head_dim=d, rope_theta=cfg.rope_theta, hidden_size=0, num_attention_heads=1), 128, DT)
    q = qkv[:, : nh * d].view(t, nh, d)
    k = qkv[:, nh * d:(nh + nkv) * d].view(t, nkv, d)
    v = qkv[:, (nh + nkv) * d:].view(t, nkv, d)
    qe, ke = od.apply_rope(q, k, oc[positions.long()], osn[positions.long()])
    assert torch.equal(q_out.cpu().view(t, nh, d), qe), "RoPE(q) must be bit-exact (same tables, same rounding points)"
    assert torch.equal(k_lin.cpu().view(t, nkv, d), ke)
    assert torch.equal(v_lin.cpu().view(t, nkv, d), v)
    kcc, vcc = kc.cpu(), vc.cpu()
    for i in range(t):
        sl = int(slot[i])
        if sl < 0:
            continue
        assert torch.equal(kcc[sl // page, :, sl % page], ke[i])
        assert torch.equal(vcc[sl // page, :, sl % page], v[i])
    # nothing else was written
    written = torch.zeros(n_pages, page, dtype=torch.bool)
    for sl in slot.tolist():
        if sl >= 0:
            written[sl // page, sl % page] = True
    assert (kcc.abs().sum((1, 3))[~written] == 0).all()


def test_embed_gather_skips_negative_ids():
    c = ctx()
    table = torch.randn(50, 64).to(DT)
    ids = torch.tensor([3, -1, 49, 0, -1], dtype=torch.int32)
    out = torch.full((5, 64), 7.0, device="cuda", dtype=DT)
    c.embed_gather(table.cuda(), ids.cuda(), out)
    o = out.cpu()
    assert torch.equal(o[0], table[3]) and torch.equal(o[2], table[49]) and torch.equal(o[3], table[0])
    assert (o[1] == 7).all() and (o[4] == 7).all()


@pytest.mark.parametrize("vocab", [1000, 152064, 777])
def test_greedy_advance(vocab):
    c = ctx()
    b, page, max_pages = 3, 16, 4
    g = torch.Generator().manual_seed(vocab)
    logits = torch.randn(b, vocab, generator=g).to(DT)
    logits[1, 5] = logits[1, 400] = 50.0            # tie: the first index wins, like torch.argmax
    ld = logits.cuda()
    out_tokens = torch.zeros(b, 8, dtype=torch.int32, device="cuda")
    step = torch.tensor([2, 0], dtype=torch.int32, device="cuda")
    cur = torch.zeros(b, dtype=torch.int32, device="cuda")
    pos = torch.tensor([14, 15, 31], dtype=torch.int32, device="cuda")
    sl = torch.tensor([15, 16, 32], dtype=torch.int32, device="cuda")
    slot = torch.zeros(b, dtype=torch.int32, device="cuda")
    pt = torch.tensor([[3, 7, 0, 0], [1, 9, 0, 0], [2, 4, 6, 0]], dtype=torch.int32, device="cuda")
    c.greedy_advance(ld, b, out_tokens, step, cur, pos, sl, slot, pt, page)
    torch.cuda.synchronize()
    ref = torch.argmax(logits.float(), dim=-1)
    assert ref[1] == 5
    assert cur.cpu().tolist() == ref.tolist()
    assert out_tokens.cpu()[:, 2].tolist() == ref.tolist() and step.cpu().tolist() == [3, 0]
    assert pos.cpu().tolist() == [15, 16, 32] and sl.cpu().tolist() == [16, 17, 33]
    assert slot.cpu().tolist() == [3 * 16 + 15, 9 * 16 + 0, 6 * 16 + 0]


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("partial", [True, False])
def test_qkv_rope_cache_qwen3_qk_norm(d, partial):
    """Qwen3 / ChatTS-8B: per-head RMSNorm of q and k before RoPE (scalar kernel on split-K partials, vector kernel otherwise)."""
    c = ctx()
    from chatts_b200.config import ChatTSConfig
    from chatts_b200.model import rope_tables
    t, nh, nkv, page = 9, 8, 2, 16
    cfg = ChatTSConfig.tiny(head_dim=d, num_attention_heads=nh, num_key_value_heads=nkv)
    width = (nh + 2 * nkv) * d
    g = torch.Generator().manual_seed(31 + d)
    qn = (torch.rand(d, generator=g) + 0.5).to(DT)
    kn = (torch.rand(d, generator=g) + 0.5).to(DT)
    positions = torch.arange(t, dtype=torch.int32) * 3
    if partial:
        s = 2
        part = torch.randn(s, t, width, generator=g)
        qkv = part.sum(0).to(DT)
        src = part.cuda()
    else:
        s = 1
        qkv = torch.randn(t, width, generator=g).to(DT)
        src = qkv.cuda()
    cos, sin = rope_tables(cfg, 64, DT, "cuda")
    kc = torch.zeros(4, nkv, page, d, device="cuda", dtype=DT)
    vc = torch.zeros_like(kc)
    slot = torch.arange(t, dtype=torch.int32) + 5
    q_out = torch.empty(t, nh * d, device="cuda", dtype=DT)
    k_lin = torch.empty(t, nkv * d, device="cuda", dtype=DT)
    v_lin = torch.empty(t, nkv * d, device="cuda", dtype=DT)
    c.qkv_rope_cache(src, partial, s, None, positions.cuda(), cos, sin, slot.cuda(), q_out, kc, vc, k_lin, v_lin, t, nh, nkv, d, page,
                     qn.cuda(), kn.cuda(), 1e-6)
    torch.cuda.synchronize()
    oc, osn = od.rope_tables(dict(head_dim=d, rope_theta=cfg.rope_theta, hidden_size=0, num_attention_heads=1), 64, DT)
    q = od.rms_norm(qkv[:, : nh * d].view(t, nh, d), qn, 1e-6)
    k = od.rms_norm(qkv[:, nh * d:(nh + nkv) * d].view(t, nkv, d), kn, 1e-6)
    v = qkv[:, (nh + nkv) * d:].view(t, nkv, d)
    qe, ke = od.apply_rope(q, k, oc[positions.long()], osn[positions.long()])
    # the norm's fp32 sum order differs (shuffle tree vs torch), so allow one dtype ulp after the exact RoPE rounding chain
    assert rel_err(q_out.view(t, nh, d), qe) <= 2 ** -7 and rel_err(k_lin.view(t, nkv, d), ke) <= 2 ** -7
    assert torch.equal(v_lin.cpu().view(t, nkv, d), v)
    kcc = kc.cpu()
    for i in range(t):
        sl = int(slot[i])
        assert torch.equal(kcc[sl // page, :, sl % page], k_lin.cpu().view(t, nkv, d)[i])
