"""Oracle: Value-Preserved Time-Series Encoder.  TEST INFRASTRUCTURE ONLY.

CPU (torch) restatement of ``TimeSeriesEmbedding`` / ``get_patch_cnt`` in
chatts/vllm/chatts_vllm.py:61-207 and of the batch assembly in :493-562
(NetManAIOps/ChatTS @ 09fae34).  Pinned by tests/golden/ts_encoder_*.npz, which were produced by
executing the reference class itself (tests/golden/make_golden.py).

Weights are passed as a plain dict with the checkpoint's names:
  ``mlp.{2*i}.weight`` [out,in], ``mlp.{2*i}.bias`` [out]   (nn.Sequential of Linear,GELU,...  :83-91)
  ``position_embedding.weight`` [max_sequence_length+1, embedding_dim]                     (:75)
"""

import torch


def input_size(cfg):
    """chatts_vllm.py:73-81."""
    p = cfg["patch_size"]
    if cfg.get("use_position_embedding", False):
        return p + cfg.get("embedding_dim", 16) * p
    if cfg.get("use_position_idx", False):
        return 2 * p
    return p


def patch_count(x, cfg):
    """chatts_vllm.py:94-100 / :198-207.  x: [N, 2L, 1] (or [N, L*F]) -> (valid_len, patch_cnt) int64."""
    n = x.shape[0]
    xr = x.reshape(n, -1, cfg["num_features"])
    valid = xr[:, :, -1].long().sum(dim=1)
    p = int(cfg["patch_size"])
    return valid, (valid + p - 1) // p


def patch_rows(x, cfg, weights):
    """chatts_vllm.py:107-183: the [sum(P), in0] matrix fed to the MLP (series order, then patch
    order :187) and patch_cnt.  Values keep x's dtype."""
    n = x.shape[0]
    p = int(cfg["patch_size"])
    xr = x.reshape(n, -1, cfg["num_features"])
    valid, pc = patch_count(x, cfg)
    use_emb = cfg.get("use_position_embedding", False)
    use_idx = cfg.get("use_position_idx", False) and not use_emb
    max_valid = int(valid.max().item()) if n > 0 else 0
    rows = []
    for i in range(n):
        vl, cnt = int(valid[i]), int(pc[i])
        if cnt == 0:                                          # :110-111
            continue
        vals = xr[i, :vl, 0]                                  # :114 first vl values (mask is a prefix)
        pad = cnt * p - vl                                    # :115-116
        if pad > 0:
            if not use_emb:
                # :128 reads self.padding_idx, defined only under use_position_embedding (:76):
                # the reference raises AttributeError here.  Mirror it.
                raise AttributeError("'TimeSeriesEmbedding' object has no attribute 'padding_idx'")
            vals = torch.cat([vals, vals[-1:].repeat(pad)])   # :121-125 pad with LAST VALID value
        vals = vals.reshape(cnt, p)                           # :132
        if use_emb:
            pos = torch.arange(vl)
            if pad > 0:                                       # :128-129 padding id = max_sequence_length
                pos = torch.cat([pos, torch.full((pad,), int(cfg["max_sequence_length"]))])
            emb = weights["position_embedding.weight"][pos.reshape(cnt, p)]   # :165  [cnt,p,E]
            rows.append(torch.cat([vals, emb.flatten(1).to(vals.dtype)], dim=1))   # :178-182
        elif use_idx:
            # :145-154  (unreachable with pad>0 -- see above -- so pos never carries the -1 filler)
            pos = torch.arange(vl, dtype=torch.float32) / max(1, max_valid - 1)
            inter = torch.stack([vals.reshape(-1), pos.to(vals.dtype)], dim=1)
            rows.append(inter.reshape(cnt, 2 * p))
        else:
            rows.append(vals)                                 # :157-158
    if rows:
        return torch.cat(rows, dim=0), pc
    return torch.empty(0, input_size(cfg), dtype=x.dtype), pc


def gelu_erf(x):
    """nn.GELU() default = exact erf form (chatts_vllm.py:87)."""
    return torch.nn.functional.gelu(x)


def mlp(rows, cfg, weights):
    """chatts_vllm.py:83-91,186-188: Linear+GELU x (n-1), then Linear.  Computes in rows.dtype the way
    nn.Linear does (fp32 accumulate inside the matmul, output rounded to dtype after the bias)."""
    h = rows
    n_layers = int(cfg["num_layers"])
    for li in range(n_layers):
        w = weights[f"mlp.{2 * li}.weight"].to(h.dtype)
        b = weights[f"mlp.{2 * li}.bias"].to(h.dtype)
        h = torch.nn.functional.linear(h, w, b)
        if li < n_layers - 1:
            h = gelu_erf(h)
    return h


def forward(x, cfg, weights):
    """TimeSeriesEmbedding.forward (chatts_vllm.py:93-193): x [N, 2L, 1] -> (feats [sum P, H], patch_cnt [N])."""
    rows, pc = patch_rows(x, cfg, weights)
    if rows.shape[0] == 0:
        return torch.empty(0, int(cfg["hidden_size"])), pc    # :191 (default dtype)
    return mlp(rows, cfg, weights), pc


def assemble_batch(series_list, dtype=torch.float16):
    """chatts_vllm.py:510-531: zero-padded [sum rows, 2Lmax, F] tensor from per-series [1, 2L_i, F] arrays.
    The reference hard-casts to float16 (:517-519,524); ``dtype`` lets the caller pick the model dtype."""
    arrs = [torch.as_tensor(a) for a in series_list]
    max_len = max(a.shape[1] for a in arrs)
    total = sum(a.shape[0] for a in arrs)
    feat = arrs[0].shape[2]
    out = torch.zeros(total, max_len, feat, dtype=dtype)
    r = 0
    for a in arrs:
        out[r:r + a.shape[0], : a.shape[1], :] = a.to(dtype)
        r += a.shape[0]
    return out


def split_by_patch_cnt(feats, patch_cnt):
    """chatts_vllm.py:545-560: list of [P_i, H] (empty tensors for P_i == 0)."""
    out, s = [], 0
    for c in patch_cnt.tolist():
        out.append(feats[s:s + c])
        s += c
    return out
