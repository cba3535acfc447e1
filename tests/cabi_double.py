"""TEST INFRASTRUCTURE ONLY: a torch/CPU test double of the C-ABI binding (`chatts_b200._cabi.Context`).

It lets the `-m "not gpu"` suite drive the HOST logic of the package (prefill / decode orchestration, paged-KV
bookkeeping, generate(), the vLLM-shaped front end) on a machine without a GPU: every entry point of the binding is
restated with plain torch ops in the same rounding discipline (fp32 accumulate, result rounded to the model dtype).
It is injected by the `cabi_double` fixture (monkeypatch of `_cabi.get_context`); the package never imports it and has
no CPU fallback of its own -- without the fixture `get_context()` raises when the CUDA library or the GPU is missing.
"""
import numpy as np
import torch
import torch.nn.functional as F


class TorchDouble:
    arch = "test-double"
    launches = 0

    def __init__(self, split=1):
        self.split = split

    def ts_patch_count(self, x, nf, p):
        n = x.shape[0]; xr = x.reshape(n, -1, nf)
        valid = xr[:, :, -1].long().sum(1).int(); cnt = (valid + p - 1) // p
        off = torch.cat([torch.zeros(1, dtype=torch.int32), cnt.cumsum(0).int()]); mx = valid.max().reshape(1).int() if n else torch.zeros(1, dtype=torch.int32)
        return valid, cnt.int(), off, mx
    def ts_patchify(self, x, nf, p, mode, pos_table, emb, maxseq, valid, off, mx, max_patches, rows):
        n = x.shape[0]; xr = x.reshape(n, -1, nf)
        for s in range(n):
            vl = int(valid[s]); cnt = (vl + p - 1) // p
            for pi in range(cnt):
                r = int(off[s]) + pi
                for j in range(p):
                    pt = pi * p + j; src = pt if pt < vl else vl - 1
                    rows[r, j] = xr[s, src, 0]
                    if mode == 1:
                        idx = pt if pt < vl else maxseq
                        rows[r, p + j * emb: p + (j + 1) * emb] = pos_table[idx]
    def suggest_split(self, n, k, t, dual=False): return self.split if k >= 64 * self.split else 1
    def gemm(self, x, w, out, *, w2=None, bias=None, residual=None, row_map=None, epilogue=0, split_k=1, t=None, splitk_ws=None, tile_counters=None,
             next_w=None, next_split=1, next_bytes=0):          # next_*: L2 prefetch hint, no effect on the result
        t = x.shape[0] if t is None else t
        xx = x[:t].float(); acc = xx @ w.float().T
        dt = x.dtype
        if epilogue == 6:
            a2 = acc.view(t, -1, 2, 64); g = a2[:, :, 0].reshape(t, -1).to(dt); u = a2[:, :, 1].reshape(t, -1).to(dt)
            out[:t] = (F.silu(g.float()).to(dt).float() * u.float()).to(dt); return
        if epilogue == 5:
            out.view(-1)[: t * w.shape[0]].view(t, w.shape[0]).copy_(acc); return
        if epilogue == 3:
            K = w.shape[1]; kb = (K + 63) // 64
            o = out.view(-1)[: split_k * t * w.shape[0]].view(split_k, t, w.shape[0])
            for s in range(split_k):
                a, b = kb * s // split_k * 64, kb * (s + 1) // split_k * 64
                o[s] = xx[:, a:b] @ w.float()[:, a:b].T
            return
        if epilogue == 2:
            g = acc.to(dt); u = (xx @ w2.float().T).to(dt)
            r = (F.silu(g.float()).to(dt).float() * u.float()).to(dt)
        else:
            r = acc + (bias.float() if bias is not None else 0)
            if epilogue == 1: r = F.gelu(r.to(dt).float())
            if epilogue == 4: r = r.to(dt).float() + residual[:t].float()
            r = r.to(dt)
        if row_map is not None:
            for i in range(t):
                if row_map[i] >= 0: out[int(row_map[i])] = r[i]
        else:
            out[:t] = r
    def reduce_bias_act(self, part, s, t, n, bias, act, out, row_map=None):
        r = part.view(-1)[: s * t * n].view(s, t, n).sum(0) + (bias.float() if bias is not None else 0)
        dt = out.dtype
        if act == 1: r = F.gelu(r.to(dt).float())
        r = r.to(dt)
        if row_map is not None:
            for i in range(t):
                if row_map[i] >= 0: out[int(row_map[i])] = r[i]
        else: out[:t] = r
    def reduce_residual_rmsnorm(self, part, s, rin, rout, w, eps, nout, t=None):
        t = rin.shape[0] if t is None else t; h = rin.shape[-1]; dt = rin.dtype
        x = rin[:t]
        if s:
            p = part.view(-1)[: s * t * h].view(s, t, h).sum(0).to(dt)
            x = (x.float() + p.float()).to(dt); rout[:t] = x
        if nout is not None:
            xf = x.float(); xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
            nout[:t] = w * xf.to(dt)
    def reduce_swiglu(self, part, s, t, inter, out, interleaved=False):
        p = part.view(-1)[: s * t * 2 * inter].view(s, t, 2 * inter).sum(0); dt = out.dtype
        if interleaved:
            pp = p.view(t, -1, 2, 64); g, u = pp[:, :, 0].reshape(t, inter), pp[:, :, 1].reshape(t, inter)
        else:
            g, u = p[:, :inter], p[:, inter:]
        out[:t] = (F.silu(g.to(dt).float()).to(dt).float() * u.to(dt).float()).to(dt)
    def qkv_rope_cache(self, src, partial, s, bias, pos, cos, sin, slot, q_out, kc, vc, k_lin, v_lin, t, nh, nkv, d, page, q_norm_w=None, k_norm_w=None, norm_eps=1e-6):
        W = (nh + 2 * nkv) * d; dt = q_out.dtype
        if partial:
            x = src.view(-1)[: s * t * W].view(s, t, W).sum(0) + (bias.float() if bias is not None else 0); x = x.to(dt)
        else: x = src[:t]
        q = x[:, : nh * d].view(t, nh, d); k = x[:, nh * d:(nh + nkv) * d].view(t, nkv, d); v = x[:, (nh + nkv) * d:].view(t, nkv, d)
        if q_norm_w is not None:
            def hn(a, w):
                af = a.float()
                return w * (af * torch.rsqrt(af.pow(2).mean(-1, keepdim=True) + norm_eps)).to(dt)
            q, k = hn(q, q_norm_w), hn(k, k_norm_w)
        c = torch.cat([cos[pos[:t].long()]] * 2, -1)[:, None]; sn = torch.cat([sin[pos[:t].long()]] * 2, -1)[:, None]
        rot = lambda a: torch.cat((-a[..., d // 2:], a[..., : d // 2]), -1)
        q = q * c + rot(q) * sn; k = k * c + rot(k) * sn
        q_out[:t] = q.reshape(t, -1)
        if k_lin is not None: k_lin[:t] = k.reshape(t, -1); v_lin[:t] = v.reshape(t, -1)
        for i in range(t if (slot is not None and kc is not None) else 0):
            sl = int(slot[i])
            if sl >= 0: kc[sl // page, :, sl % page] = k[i]; vc[sl // page, :, sl % page] = v[i]
    def embed_gather(self, table, ids, out, t=None):
        t = ids.shape[0] if t is None else t
        for i in range(t):
            if ids[i] >= 0: out[i] = table[int(ids[i])]
    def _attn(self, q, k, v, nrep, p0):
        T, nh, d = q.shape; S = k.shape[0]
        kk = k.repeat_interleave(nrep, 1).float(); vv = v.repeat_interleave(nrep, 1).float()
        w = torch.einsum("thd,shd->hts", q.float(), kk) * d ** -0.5
        m = torch.arange(S)[None] > (torch.arange(T)[:, None] + p0)
        w = w.masked_fill(m[None], -1e30).softmax(-1)
        return torch.einsum("hts,shd->thd", w, vv).reshape(T, nh * d)
    def attn_prefill(self, q, k, v, cu, B, maxlen, nh, nkv, d, scale, out):
        for b in range(B):
            a, e = int(cu[b]), int(cu[b + 1])
            out[a:e] = self._attn(q[a:e].view(-1, nh, d), k[a:e].view(-1, nkv, d), v[a:e].view(-1, nkv, d), nh // nkv, 0).to(out.dtype)
    def attn_decode_workspace_floats(self, *a): return 8
    def attn_decode(self, q, kc, vc, pt, sl, B, nh, nkv, d, page, scale, splits, ws, out):
        for b in range(B):
            n = int(sl[b]); pages = pt[b, : (n + page - 1) // page].long()
            kk = kc[pages].permute(0, 2, 1, 3).reshape(-1, nkv, d)[:n]; vv = vc[pages].permute(0, 2, 1, 3).reshape(-1, nkv, d)[:n]
            out[b] = self._attn(q[b:b + 1].view(1, nh, d), kk, vv, nh // nkv, n - 1).to(out.dtype)[0]
    def greedy_advance(self, logits, B, out_tokens, step_ptr, cur, pos, sl, slot, pt, page):
        tok = logits[:B].float().argmax(-1).int(); st = int(step_ptr[0])
        out_tokens[:B, st] = tok; cur[:B] = tok; pos[:B] += 1; sl[:B] += 1
        for b in range(B):
            slot[b] = pt[b, int(pos[b]) // page] * page + int(pos[b]) % page
        step_ptr += 1

    # ------------------------------------------------------------------ A9: LoRA fine-tune step (include/chatts_b200.h "A9")
    def _attn_lse(self, q, k, v, nrep):
        T, nh, d = q.shape
        kk = k.repeat_interleave(nrep, 1); vv = v.repeat_interleave(nrep, 1)
        w = torch.einsum("thd,shd->hts", q, kk) * d ** -0.5
        m = torch.arange(T)[None] > torch.arange(T)[:, None]
        w = w.masked_fill(m[None], float("-inf"))
        return torch.einsum("hts,shd->thd", w.softmax(-1), vv).reshape(T, nh * d), torch.logsumexp(w, -1).T      # [T, nh]
    def attn_prefill_lse(self, q, k, v, cu, B, maxlen, nh, nkv, d, scale, out, lse):
        for b in range(B):
            a, e = int(cu[b]), int(cu[b + 1])
            o, l = self._attn_lse(q[a:e].view(-1, nh, d).float(), k[a:e].view(-1, nkv, d).float(), v[a:e].view(-1, nkv, d).float(), nh // nkv)
            out[a:e] = o.to(out.dtype); lse[a:e] = l
    def attn_bwd(self, q, k, v, out, dout, lse, cu, B, maxlen, nh, nkv, d, scale, delta_ws, dq, dk, dv):
        for b in range(B):
            a, e = int(cu[b]), int(cu[b + 1])
            with torch.enable_grad():
                qq = q[a:e].view(-1, nh, d).float().requires_grad_(True); kk = k[a:e].view(-1, nkv, d).float().requires_grad_(True)
                vv = v[a:e].view(-1, nkv, d).float().requires_grad_(True)
                o, _ = self._attn_lse(qq, kk, vv, nh // nkv)
                gq, gk, gv = torch.autograd.grad(o, [qq, kk, vv], dout[a:e].float())
            dq[a:e] = gq.reshape(e - a, -1).to(dq.dtype); dk[a:e] = gk.reshape(e - a, -1).to(dk.dtype); dv[a:e] = gv.reshape(e - a, -1).to(dv.dtype)
    @staticmethod
    def _gu_split(gu, t, inter, interleaved):
        x = gu[:t]
        if interleaved:
            xx = x.view(t, -1, 2, 64); return xx[:, :, 0].reshape(t, inter), xx[:, :, 1].reshape(t, inter)
        return x[:, :inter], x[:, inter:]
    def swiglu(self, gu, t, inter, out, interleaved=True):
        g, u = self._gu_split(gu, t, inter, interleaved); dt = out.dtype
        out[:t] = (F.silu(g.float()).to(dt).float() * u.float()).to(dt)
    def swiglu_bwd(self, gu, dact, t, inter, dgu, interleaved=True):
        g, u = self._gu_split(gu, t, inter, interleaved); g, u, d = g.float(), u.float(), dact[:t].float(); dt = dgu.dtype
        sg = torch.sigmoid(g)
        dg = (d * u * sg * (1 + g * (1 - sg))).to(dt); du = (d * F.silu(g).to(dt).float()).to(dt)
        if interleaved:
            o = dgu[:t].view(t, -1, 2, 64); o[:, :, 0] = dg.view(t, -1, 64); o[:, :, 1] = du.view(t, -1, 64)
        else:
            dgu[:t, :inter] = dg; dgu[:t, inter:] = du
    def rmsnorm_bwd(self, dy, x, w, eps, dres_in, dx_out, t=None):
        t = x.shape[0] if t is None else t
        with torch.enable_grad():
            xx = x[:t].float().requires_grad_(True)
            y = w.float() * (xx * torch.rsqrt(xx.pow(2).mean(-1, keepdim=True) + eps))
            (gx,) = torch.autograd.grad(y, [xx], dy[:t].float())
        r = gx + (dres_in[:t].float() if dres_in is not None else 0)
        dx_out[:t] = r.to(dx_out.dtype)
    def qkv_rope_bwd(self, dq, dk, dv, qkv, pos, cos, sin, q_norm_w, k_norm_w, norm_eps, dqkv, t, nh, nkv, d):
        c = torch.cat([cos[pos[:t].long()]] * 2, -1)[:, None].float(); sn = torch.cat([sin[pos[:t].long()]] * 2, -1)[:, None].float()
        rot = lambda a: torch.cat((-a[..., d // 2:], a[..., : d // 2]), -1)
        with torch.enable_grad():
            x = qkv[:t].float().requires_grad_(True)
            q = x[:, : nh * d].view(t, nh, d); k = x[:, nh * d:(nh + nkv) * d].view(t, nkv, d); v = x[:, (nh + nkv) * d:]
            if q_norm_w is not None:
                hn = lambda a, w: w.float() * (a * torch.rsqrt(a.pow(2).mean(-1, keepdim=True) + norm_eps))
                q, k = hn(q, q_norm_w), hn(k, k_norm_w)
            q = q * c + rot(q) * sn; k = k * c + rot(k) * sn
            tot = (q.reshape(t, -1) * dq[:t].float()).sum() + (k.reshape(t, -1) * dk[:t].float()).sum() + (v * dv[:t].float()).sum()
            (gx,) = torch.autograd.grad(tot, [x])
        dqkv[:t] = gx.to(dqkv.dtype)
    def ce_loss_grad(self, logits, targets, n, grad_scale, row_loss, loss_out, accumulate=False):
        lg = logits[:n].float(); tg = targets[:n].long()
        lse = torch.logsumexp(lg, -1)
        row_loss[:n] = lse - lg.gather(1, tg[:, None])[:, 0]
        p = torch.softmax(lg, -1); p[torch.arange(n), tg] -= 1
        logits[:n] = (p * grad_scale).to(logits.dtype)
        loss_out[0] = (loss_out[0] if accumulate else 0) + grad_scale * row_loss[:n].sum()
    def gather_rows(self, src, idx, n, dst):
        ii = idx[:n].long()
        dst[:n] = torch.where((ii >= 0)[:, None], src[ii.clamp(min=0)], torch.zeros((), dtype=src.dtype))
    def lora_wgrad(self, p, p_col0, p_il, m, q, q_col0, r, t, scale, out, so_m, so_r):
        i = torch.arange(m)
        cols = p_col0 + i if p_il == 0 else (i // 64) * 128 + i % 64 + (64 if p_il == 2 else 0)
        g = scale * (p[:t][:, cols].float().T @ q[:t, q_col0:q_col0 + r].float())                # [m, r]
        torch.as_strided(out, (m, r), (so_m, so_r)).add_(g)
    def adamw(self, p, g, m, v, lr, b1, b2, eps, wd, step, grad_scale=None):
        gg = g * (float(grad_scale[0]) if grad_scale is not None else 1.0)
        p.mul_(1 - lr * wd); m.mul_(b1).add_(gg, alpha=1 - b1); v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
        p.sub_((lr / (1 - b1 ** step)) * m / (v.sqrt() / (1 - b2 ** step) ** 0.5 + eps))
    def grad_norm_ws_floats(self): return 8
    def grad_norm_clip(self, g, max_norm, ws, out):
        n = g.norm(); out[0] = n; out[1] = min(1.0, max_norm / (float(n) + 1e-6)) if max_norm > 0 else 1.0
    def lora_pack(self, master, desc, n_desc, max_elems, work):
        import numpy as np
        D = desc.view(n_desc, -1).tolist()
        for src_off, rows, cols, dst_off, dst_ld, row0, il, col0, dstT_off, dstT_ld, sbits, _ in D:
            scale = float(np.uint32(sbits).view(np.float32))
            src = (master[src_off: src_off + rows * cols].view(rows, cols) * scale).to(work.dtype)
            i = torch.arange(rows)
            rm = row0 + i if il == 0 else (i // 64) * 128 + i % 64 + (64 if il == 2 else 0)
            for a in range(rows):
                r0 = int(rm[a])
                work[dst_off + r0 * dst_ld + col0: dst_off + r0 * dst_ld + col0 + cols] = src[a]
                work[dstT_off + col0 * dstT_ld + r0: dstT_off + (col0 + cols) * dstT_ld + r0: dstT_ld] = src[a]

    # ------------------------------------------------------------------ K13 sampled variant (csrc/sampling.cu)
    @staticmethod
    def _splitmix64(x):
        M = (1 << 64) - 1
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)
    @classmethod
    def sample_reference(cls, row, temperature, top_k, top_p, seed, step, b):
        """(token, cdf float64 over the kept set in index order, target, kept mask) -- the statement cts_sample_advance follows."""
        import numpy as np
        x = row.float().numpy(); V = x.shape[0]
        bits = row.contiguous().view(torch.int16).numpy().astype(np.uint16).astype(np.uint32)
        key = np.where(bits & 0x8000, (~bits) & 0xFFFF, bits | 0x8000)
        ok = ~np.isnan(x)
        m = x[ok].max()
        p = np.where(ok, np.exp((x - m) * np.float32(1.0 / temperature), dtype=np.float32), np.float32(0))
        kmax = int(key[ok].max()); tau = 0
        if top_k and 0 < top_k < V:
            lo, hi = 0, kmax + 1
            while hi - lo > 1:
                mid = lo + ((hi - lo) >> 1)
                lo, hi = (mid, hi) if int((ok & (key >= mid)).sum()) >= top_k else (lo, mid)
            tau = lo
        mass = float(p[ok & (key >= tau)].sum(dtype=np.float64))
        if top_p is not None and 0.0 < top_p < 1.0:
            lo, hi, target = tau, kmax + 1, top_p * mass
            while hi - lo > 1:
                mid = lo + ((hi - lo) >> 1)
                lo, hi = (mid, hi) if float(p[ok & (key >= mid)].sum(dtype=np.float64)) >= target else (lo, mid)
            tau = lo
        kept = ok & (key >= tau)
        cdf = np.cumsum(np.where(kept, p, 0), dtype=np.float64)
        M = (1 << 64) - 1
        h = cls._splitmix64((int(seed) & M) ^ cls._splitmix64(((int(step) & 0xFFFFFFFF) << 32) | (int(b) & 0xFFFFFFFF)))
        u = float(h >> 40) / 16777216.0
        target = u * cdf[-1]
        tok = int(np.searchsorted(cdf, target, side="right"))
        tok = min(tok, int(np.nonzero(kept)[0][-1]))
        return tok, cdf, target, kept
    def sample_advance(self, logits, B, temperature, top_k, top_p, seed, out_tokens, step_ptr, cur, pos, sl, slot, pt, page):
        st = int(step_ptr[0])
        for b in range(B):
            tok = self.sample_reference(logits[b], temperature, top_k, top_p, seed, st, b)[0]
            out_tokens[b, st] = tok
            if cur is not None: cur[b] = tok
            if pos is not None:
                pos[b] += 1
                if sl is not None: sl[b] += 1
                if slot is not None and pt is not None: slot[b] = pt[b, int(pos[b]) // page] * page + int(pos[b]) % page
        step_ptr[0] += 1

    # ------------------------------------------------------------------ W4A16 decode GEMM (csrc/gemm_w4.cu)
    def gemm_w4_suggest_split(self, n, k): return self.split if k >= 128 * self.split else 1
    def gemm_w4(self, x, qw, scales, zeros, group_size, out, split_k, t=None):
        from chatts_b200.weights import dequantize_w4
        w = dequantize_w4(qw, scales, zeros, group_size)
        self.gemm(x, w, out, epilogue=3, split_k=split_k, t=t)

    def gemm_w4_mma_suggest_split(self, n, k, t=1): return self.split if k >= 256 * self.split else 1
    def gemm_w4_mma(self, x, qwf, szp, n, group_size, out, split_k, t=None):
        """The fragment-major layout decoded back to a dense weight (the inverse of weights.py:repack_w4_mma, written independently)."""
        k = szp.shape[1] * int(group_size)
        assert k % 128 == 0 and (int(group_size) == 64 or int(group_size) % 128 == 0) and 1 <= split_k <= k // 128      # cts_gemm_w4f_args
        key = (qwf.data_ptr(), szp.data_ptr(), int(n), int(group_size), x.dtype)
        cache = self.__dict__.setdefault("_w4f_cache", {})
        if key in cache:                                     # the weights of a model are static: decode the layout once per tensor
            self.gemm(x, cache[key], out, epilogue=3, split_k=split_k, t=t)
            return
        tiles = szp.shape[0]
        b = qwf.view(tiles, k // 64, 16, 32, 4, 4).to(torch.int64)
        word = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16) | (b[..., 3] << 24)           # [tile, kb, m, lane, ks]
        q = torch.zeros(tiles * 256, k, dtype=torch.int64)
        for m in range(16):
            for lane in range(32):
                g, tq = lane >> 2, lane & 3
                for nib in range(8):
                    row = m * 16 + g + 8 * (nib & 1)
                    kk = 2 * tq + 8 * ((nib >> 1) & 1) + (nib >> 2)
                    v = (word[:, :, m, lane, :] >> (4 * nib)) & 0xF                         # [tile, kb, ks]
                    for ks in range(4):
                        q.view(tiles, 256, k // 64, 64)[:, row, :, 16 * ks + kk] = v[:, :, ks]
        u = szp.to(torch.int64) & 0xFFFFFFFF                                                   # [tile, group, 256]
        sc = (u & 0xFFFF).to(torch.int16).view(x.dtype).permute(0, 2, 1).reshape(tiles * 256, -1)
        magic = 0x4300 if x.dtype == torch.bfloat16 else 0x6400
        zp = ((u >> 16) - magic).permute(0, 2, 1).reshape(tiles * 256, -1)
        grp = torch.arange(k) // int(group_size)
        w = (sc.to(torch.float32)[:, grp] * (q - zp[:, grp]).to(torch.float32)).to(x.dtype)[:n]
        cache[key] = w
        self.gemm(x, w, out, epilogue=3, split_k=split_k, t=t)

    # ------------------------------------------------------------------ repetition penalty (csrc/sampling.cu)
    def rep_penalty_mark(self, tokens, rows, seen, vocab):
        tk = tokens.reshape(-1).tolist()
        rw = rows.reshape(-1).tolist() if rows is not None else list(range(len(tk)))
        u = seen.numpy().view(np.uint32)                     # same memory: bits set in place
        for t, r in zip(tk, rw):
            if 0 <= t < vocab and r >= 0:
                u[r, t >> 5] |= np.uint32(1 << (t & 31))

    def rep_penalty_apply(self, logits, batch, seen, penalty):
        V = logits.shape[-1]
        bits = seen[:batch].to(torch.int64) & 0xFFFFFFFF
        mask = ((bits[:, :, None] >> torch.arange(32)[None, None, :]) & 1).reshape(batch, -1)[:, :V].bool()
        lg = logits[:batch].float()
        logits[:batch] = torch.where(mask, torch.where(lg > 0, lg / penalty, lg * penalty), lg).to(logits.dtype)

    # ------------------------------------------------------------------ host executors (csrc/decoder_step.cu)
    def rmsnorm(self, x, w, eps, out, t=None): self.reduce_residual_rmsnorm(None, 0, x, None, w, eps, out, t=t)
    def lm_head(self, hidden, w, logits, t=None): self.gemm(hidden, w, logits, t=t)
    def decoder_step(self, *, layers, embed, final_norm, lm_head, cos, sin, hidden, inter, nh, nkv, head_dim, eps, page_size, batch, splits,
                     attn_splits, cur_ids, positions, seq_lens, slot_map, page_table, out_tokens, step_ptr, h, xn, q, ao, act, logits, ws,
                     attn_ws, sample=True):
        """Same call sequence as cts_decoder_step: every projection through the split-K partial path."""
        T, H, I, d = batch, hidden, inter, head_dim; sq, so, sg, sd_ = splits
        self.embed_gather(embed, cur_ids, h, t=T)
        self.reduce_residual_rmsnorm(None, 0, h, None, layers[0]["ln1"], eps, xn, t=T)
        for l, w in enumerate(layers):
            self.gemm(xn, w["wqkv"], ws, epilogue=3, split_k=sq, t=T)
            self.qkv_rope_cache(ws, True, sq, w["bqkv"], positions, cos, sin, slot_map, q, w["k_cache"], w["v_cache"], None, None, T, nh, nkv, d,
                                page_size, w["q_norm"], w["k_norm"], eps)
            self.attn_decode(q, w["k_cache"], w["v_cache"], page_table, seq_lens, T, nh, nkv, d, page_size, d ** -0.5, attn_splits, attn_ws, ao)
            self.gemm(ao, w["wo"], ws, epilogue=3, split_k=so, t=T)
            self.reduce_residual_rmsnorm(ws, so, h, h, w["ln2"], eps, xn, t=T)
            self.gemm(xn, w["wgu"], ws, epilogue=3, split_k=sg, t=T)
            self.reduce_swiglu(ws, sg, T, I, act, interleaved=True)
            self.gemm(act, w["wd"], ws, epilogue=3, split_k=sd_, t=T)
            self.reduce_residual_rmsnorm(ws, sd_, h, h, layers[l + 1]["ln1"] if l + 1 < len(layers) else final_norm, eps, xn, t=T)
        self.gemm(xn, lm_head, logits, t=T)
        if sample:
            self.greedy_advance(logits, T, out_tokens, step_ptr, cur_ids, positions, seq_lens, slot_map, page_table, page_size)
    def ts_encode(self, x, nf, p, mode, pos_table, emb, maxseq, weights, biases, total_rows, out, row_map=None):
        """Same call sequence as cts_ts_encode."""
        n = x.shape[0]; xx = x.reshape(n, -1).contiguous()
        valid, cnt, off, mx = self.ts_patch_count(xx, nf, p)
        if total_rows == 0: return valid, cnt, off
        in0 = weights[0].shape[1]; H = weights[0].shape[0]
        rows = torch.zeros(total_rows, in0, dtype=xx.dtype)
        self.ts_patchify(xx, nf, p, mode, pos_table, emb, maxseq, valid, off, mx, (xx.shape[1] // nf + p - 1) // p, rows)
        h = rows
        for li, (w, b) in enumerate(zip(weights, biases)):
            last = li == len(weights) - 1
            dst = out if last else torch.empty(total_rows, H, dtype=xx.dtype)
            self.gemm(h, w, dst, bias=b, row_map=row_map if last else None, epilogue=0 if last else 1)
            h = dst
        return valid, cnt, off

    # ------------------------------------------------------------------ csrc/gemm_decode_fused.cu
    def gemm_decode_fused(self, x, w, mode, split_k, t, *, bias=None, h=None, act=None, positions=None, cos=None, sin=None, slot_map=None,
                          q_out=None, k_cache=None, v_cache=None, q_norm=None, k_norm=None, eps=1e-6, nh=0, nkv=0, head_dim=0, page_size=0,
                          norm_h=None, norm_w=None, ssq_in=None, norm_eps=1e-6, ssq_out=None, peer=None):
        """cts_gemm(CTS_EPI_PARTIAL_F32) + the matching reduce, i.e. what the cluster kernel fuses (same split order)."""
        assert peer is None, "the in-kernel all-reduce needs peer memory: not part of the single-process double"
        n = w.shape[0]
        if norm_h is not None:            # token operand = RMSNorm(norm_h) from the per-tile sums of squares
            hh = norm_h[:t].float(); rstd = torch.rsqrt(ssq_in[:t].sum(-1, keepdim=True) / hh.shape[-1] + norm_eps)
            x = norm_w * (hh * rstd).to(norm_h.dtype)
        ws = torch.empty(split_k * t * n, dtype=torch.float32)
        self.gemm(x, w, ws, epilogue=3, split_k=split_k, t=t)
        if mode == 0:
            self.reduce_residual_rmsnorm(ws, split_k, h, h, None, eps, None, t=t)
            if ssq_out is not None:
                hf = h[:t].float(); pad = (-hf.shape[1]) % 128
                ssq_out[:t] = torch.nn.functional.pad(hf, (0, pad)).view(t, -1, 128).pow(2).sum(-1)
        elif mode == 1: self.reduce_swiglu(ws, split_k, t, n // 2, act, interleaved=True)
        else: self.qkv_rope_cache(ws, True, split_k, bias, positions, cos, sin, slot_map, q_out, k_cache, v_cache, None, None, t, nh, nkv, head_dim,
                                  page_size, q_norm, k_norm, eps)
