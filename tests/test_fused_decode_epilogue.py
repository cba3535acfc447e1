"""Epilogue-level statement of csrc/gemm_decode_fused.cu on CPU: token ownership inside the cluster (t = split, split + S, ... in
groups of four), split-order summation, and the three tile-local tails (residual; SwiGLU on the interleaved gate/up tile; bias +
q/k RMSNorm + RoPE partner exchange + q_out / paged KV write with head = f / d) written thread by thread in numpy from the
kernel's index arithmetic, against the two-launch path it replaces (the torch double's gemm(partial) + reduce).  Pins the
arithmetic of part B; the mainloop is the GPU-validated one of gemm_tcgen05.cu.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import pytest
import torch

from tests.cabi_double import TorchDouble

DT = torch.bfloat16
DBL = TorchDouble()
KBM = 128


def rnd(x):
    return torch.tensor(x, dtype=torch.float32).to(DT).float().numpy()


def _partials(x, w, S):
    """[S, T, N] fp32 partial products over the K blocks of every split (as the double's gemm epilogue 3)."""
    T, N = x.shape[0], w.shape[0]
    ws = torch.empty(S * T * N, dtype=torch.float32)
    DBL.gemm(x, w, ws, epilogue=3, split_k=S, t=T)
    return ws.view(S, T, N).numpy(), ws


def _emulate(part, mode, S, T, N, *, h=None, act=None, bias=None, pos=None, cos=None, sin=None, slot=None, q_out=None, kc=None, vc=None,
             qn=None, kn=None, eps=1e-6, nh=0, nkv=0, d=0, page=0):
    tiles = (N + KBM - 1) // KBM
    owners = np.zeros(T, dtype=int)
    for tile in range(tiles):
        f0 = tile * KBM
        for split in range(S):                               # CTA (tile, split): tokens split, split + S, ... in groups of 4
            tb = split
            while tb < T:
                for u in range(4):
                    t = tb + u * S
                    if t >= T:
                        break
                    if tile == 0:
                        owners[t] += 1
                    acc = np.zeros(KBM, dtype=np.float32)
                    for ft in range(KBM):
                        if f0 + ft < N:
                            a = np.float32(0)
                            for s2 in range(S):              # split order
                                a = np.float32(a + part[s2, t, f0 + ft])
                            acc[ft] = a
                    if mode == 0:
                        for ft in range(KBM):
                            f = f0 + ft
                            if f < N:
                                h[t, f] = rnd(h[t, f] + rnd(acc[ft]))
                    elif mode == 1:
                        xch = rnd(acc[64:])
                        for ft in range(64):
                            if f0 + ft < N:
                                g = rnd(acc[ft])
                                act[t, (f0 >> 1) + ft] = rnd(rnd(g / (1 + np.exp(-g))) * xch[ft])
                    else:
                        half = d // 2
                        x = np.zeros(KBM, dtype=np.float32)
                        for ft in range(KBM):
                            f = f0 + ft
                            if f < N:
                                x[ft] = rnd(acc[ft] + (bias[f] if bias is not None else 0))
                        if qn is not None:
                            red = [float(np.sum(np.where(np.arange(32 * qq, 32 * qq + 32) + f0 < N, x[32 * qq: 32 * qq + 32] ** 2, 0)))
                                   for qq in range(4)]
                            for ft in range(KBM):
                                f = f0 + ft
                                if f >= N:
                                    continue
                                head, i = f // d, f % d
                                tot = sum(red) if d == 128 else (red[0] + red[1] if ft < 64 else red[2] + red[3])
                                nw = qn if head < nh else (kn if head < nh + nkv else None)
                                if nw is not None:
                                    inv = np.float32(1.0 / np.sqrt(np.float32(tot) / d + eps))
                                    x[ft] = rnd(nw[i] * rnd(x[ft] * inv))
                        xch = x.copy()
                        for ft in range(KBM):
                            f = f0 + ft
                            if f >= N:
                                continue
                            head, i = f // d, f % d
                            is_q, is_k = head < nh, nh <= head < nh + nkv
                            v = x[ft]
                            if is_q or is_k:
                                ii = i if i < half else i - half
                                c, sn = cos[pos[t], ii], sin[pos[t], ii]
                                xp = xch[ft + half] if i < half else xch[ft - half]
                                v = rnd(rnd(v * c) + rnd(-xp * sn)) if i < half else rnd(rnd(v * c) + rnd(xp * sn))
                            if is_q:
                                q_out[t, head * d + i] = v
                            else:
                                kvh = head - nh if is_k else head - nh - nkv
                                cache = kc if is_k else vc
                                if slot[t] >= 0:
                                    cache[slot[t] // page, kvh, slot[t] % page, i] = v
                tb += 4 * S
    return owners


@pytest.mark.parametrize("T,S", [(1, 3), (5, 2), (9, 4), (32, 7), (17, 8)])
def test_every_token_has_exactly_one_owner_per_tile(T, S):
    part = np.zeros((S, T, KBM), dtype=np.float32)
    owners = _emulate(part, 0, S, T, KBM, h=np.zeros((T, KBM), dtype=np.float32))
    assert (owners == 1).all()


def test_residual_and_swiglu_tails():
    g = torch.Generator().manual_seed(0)
    T, K, S = 9, 256, 3
    x = (torch.randn(T, K, generator=g) * 0.5).to(DT)
    w = (torch.randn(200, K, generator=g) * 0.1).to(DT)                     # ragged N: the last tile is partly empty
    part, ws = _partials(x, w, S)
    h0 = (torch.randn(T, 200, generator=g) * 0.5).to(DT)
    ref = h0.clone()
    DBL.reduce_residual_rmsnorm(ws, S, ref, ref, None, 1e-6, None, t=T)
    h = h0.float().numpy().copy()
    _emulate(part, 0, S, T, 200, h=h)
    assert np.array_equal(h, ref.float().numpy())
    wg = (torch.randn(256, K, generator=g) * 0.1).to(DT)                    # inter = 128, interleaved [2 tiles x (64 gate | 64 up)]
    part, ws = _partials(x, wg, S)
    ref = torch.empty(T, 128, dtype=DT)
    DBL.reduce_swiglu(ws, S, T, 128, ref, interleaved=True)
    act = np.zeros((T, 128), dtype=np.float32)
    _emulate(part, 1, S, T, 256, act=act)
    assert np.abs(act - ref.float().numpy()).max() <= 2 ** -8 * np.abs(ref.float().numpy()).max()      # numpy exp vs torch silu: 1 ulp


@pytest.mark.parametrize("d,nh,nkv,qk,bias", [(128, 3, 1, False, True), (64, 3, 1, True, False), (64, 4, 2, False, True), (128, 2, 2, True, False)])
def test_qkv_rope_tail(d, nh, nkv, qk, bias):
    g = torch.Generator().manual_seed(d + nh)
    T, K, S, page, pages = 7, 128, 2, 16, 4
    N = (nh + 2 * nkv) * d
    x = (torch.randn(T, K, generator=g) * 0.5).to(DT)
    w = (torch.randn(N, K, generator=g) * 0.1).to(DT)
    b = (torch.randn(N, generator=g) * 0.2).to(DT) if bias else None
    qn = (torch.rand(d, generator=g) + 0.5).to(DT) if qk else None
    kn = (torch.rand(d, generator=g) + 0.5).to(DT) if qk else None
    pos = torch.randint(0, 50, (T,), generator=g).to(torch.int32)
    ang = torch.rand(64, d // 2, generator=g) * 6.28
    cos, sin = ang.cos().to(DT), ang.sin().to(DT)
    slot = torch.randperm(pages * page, generator=g)[:T].to(torch.int32)
    slot[2] = -1
    part, ws = _partials(x, w, S)
    q_ref = torch.zeros(T, nh * d, dtype=DT)
    kc_ref, vc_ref = torch.zeros(pages, nkv, page, d, dtype=DT), torch.zeros(pages, nkv, page, d, dtype=DT)
    DBL.qkv_rope_cache(ws, True, S, b, pos, cos, sin, slot, q_ref, kc_ref, vc_ref, None, None, T, nh, nkv, d, page, qn, kn, 1e-6)
    q = np.zeros((T, nh * d), dtype=np.float32)
    kc, vc = np.zeros((pages, nkv, page, d), dtype=np.float32), np.zeros((pages, nkv, page, d), dtype=np.float32)
    f32 = lambda t_: None if t_ is None else t_.float().numpy()
    _emulate(part, 2, S, T, N, bias=f32(b), pos=pos.numpy(), cos=f32(cos), sin=f32(sin), slot=slot.numpy(), q_out=q, kc=kc, vc=vc, qn=f32(qn),
             kn=f32(kn), nh=nh, nkv=nkv, d=d, page=page)
    tol = 2 ** -7 if qk else 0.0                                            # the norm statistic is summed in another order
    for got, want in ((q, q_ref), (kc, kc_ref), (vc, vc_ref)):
        wf = want.float().numpy()
        assert np.abs(got - wf).max() <= tol * max(np.abs(wf).max(), 1e-6), (d, nh, nkv, qk)
