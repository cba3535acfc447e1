"""TS encoder on the GPU against the golden fixtures produced by the reference and against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ts_encoder as ote
from tests.gpu_util import record, rel_err

pytestmark = pytest.mark.gpu


def _load(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"ts_encoder_{tag}.npz"))
    cfg = {k[4:]: g[k].item() for k in g.files if k.startswith("cfg.")}
    w = {"ts_encoder." + k[2:]: torch.tensor(g[k]) for k in g.files if k.startswith("w.")}
    return g, cfg, w


@pytest.mark.parametrize("tag", ["posemb", "posidx", "plain", "posemb_p8"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_against_reference_fixture(golden_dir, tag, dtype):
    from chatts_b200.ts_encoder import TimeSeriesEmbedding, get_patch_cnt
    g, cfg, w = _load(golden_dir, tag)
    enc = TimeSeriesEmbedding(cfg, w, dtype=dtype)
    x = torch.tensor(g["x"])
    feats, pc = enc(x.cuda())
    assert pc.cpu().tolist() == g["patch_cnt"].tolist()                 # bit-exact integers
    assert get_patch_cnt(x.cuda().to(dtype), cfg).cpu().tolist() == g["patch_cnt"].tolist()
    # same-dtype oracle (weights and series rounded to dtype, like a half-precision checkpoint)
    wo = {k[len("ts_encoder."):]: v.to(dtype) for k, v in w.items()}
    ref, _ = ote.forward(x.to(dtype), cfg, wo)
    e_same = rel_err(feats, ref)
    e_fp32 = rel_err(feats, torch.tensor(g["feats"]))
    record("ts_encoder_fixture", tag=tag, dtype=str(dtype), err_vs_same_dtype_oracle=e_same, err_vs_reference_fp32=e_fp32)
    assert feats.shape == ref.shape
    assert e_same < 1e-2
    assert e_fp32 < 3e-2


def test_patch_rows_bit_exact(golden_dir):
    """The front end (A3-A5) moves values without arithmetic: patch rows must equal the oracle's bit for bit."""
    from chatts_b200.ts_encoder import TimeSeriesEmbedding
    g, cfg, w = _load(golden_dir, "posemb")
    dtype = torch.bfloat16
    enc = TimeSeriesEmbedding(cfg, w, dtype=dtype)
    x = torch.tensor(g["x"]).to(dtype)
    xx, valid, cnt, off, mx = enc.patch_counts(x.cuda())
    total = int(cnt.sum())
    rows = torch.empty(total, enc.input_size, device="cuda", dtype=dtype)
    enc.ctx.ts_patchify(xx, enc.num_features, enc.patch_size, enc.mode, enc.pos_table, enc.embedding_dim,
                        enc.max_sequence_length, valid, off, mx, x.shape[1] // 2 // enc.patch_size + 1, rows)
    wo = {k[len("ts_encoder."):]: v.to(dtype) for k, v in w.items()}
    ref_rows, pc = ote.patch_rows(x, cfg, wo)
    assert torch.equal(rows.cpu(), ref_rows)
    assert valid.cpu().tolist() == g["lengths"].tolist()
    assert off.cpu().tolist() == np.concatenate([[0], np.cumsum(g["patch_cnt"])]).tolist()
    assert int(mx) == int(g["lengths"].max())


def test_ragged_without_posemb_raises(golden_dir):
    from chatts_b200.ts_encoder import TimeSeriesEmbedding
    g, cfg, w = _load(golden_dir, "plain")
    enc = TimeSeriesEmbedding(cfg, w)
    x = torch.zeros(1, 34, 1)
    x[0, :, 0] = 1.0
    with pytest.raises(AttributeError):
        enc(x.cuda())


def test_all_empty_and_scatter(golden_dir):
    from chatts_b200.ts_encoder import TimeSeriesEmbedding
    g, cfg, w = _load(golden_dir, "posemb")
    enc = TimeSeriesEmbedding(cfg, w)
    feats, pc = enc(torch.zeros(3, 64, 1).cuda())
    assert feats.shape == (0, cfg["hidden_size"]) and pc.tolist() == [0, 0, 0]
    # scatter into a bigger buffer through row_map
    x = torch.tensor(g["x"]).cuda()
    feats, pc = enc(x)
    n = feats.shape[0]
    big = torch.zeros(2 * n, cfg["hidden_size"], device="cuda", dtype=torch.bfloat16)
    rmap = (torch.arange(n, dtype=torch.int32) * 2 + 1).cuda()
    enc.encode(x, out=big, row_map=rmap)
    assert torch.equal(big[1::2], feats) and (big[0::2] == 0).all()


def test_large_batch_many_series():
    """cfg-3/4 like sizes: hundreds of series, variable length 64..1024; property check: row count and order."""
    from chatts_b200.ts_encoder import TimeSeriesEmbedding
    from chatts_b200.config import ChatTSConfig
    from chatts_b200.weights import synthetic_state_dict
    cfg = ChatTSConfig.tiny()
    w = {k: v for k, v in synthetic_state_dict(cfg, seed=5, dtype=torch.float32).items() if k.startswith("ts_encoder.")}
    enc = TimeSeriesEmbedding(cfg.ts, w)
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 513, size=300)
    x = torch.zeros(300, 2 * 512, 1)
    for i, n in enumerate(lens):
        x[i, 0:2 * n:2, 0] = torch.randn(int(n))
        x[i, 1:2 * n:2, 0] = 1.0
    feats, pc = enc(x.cuda())
    assert pc.cpu().tolist() == [int((n + 15) // 16) for n in lens]
    wo = {k[len("ts_encoder."):]: v.to(torch.bfloat16) for k, v in w.items()}
    ref, _ = ote.forward(x.to(torch.bfloat16), cfg.ts, wo)
    e = rel_err(feats, ref)
    record("ts_encoder_300_series", err=e)
    assert e < 1e-2


# --------------------------------------------------------------------------------------------------------------------
# The fused encoder (csrc/ts_encoder_fused.cu: patchify + every MLP layer + the row scatter in ONE launch, <= 256 patch rows)
# --------------------------------------------------------------------------------------------------------------------
def _series_batch(lens, seed=0):
    from chatts_b200.processor import sp_encoding
    rng = np.random.default_rng(seed)
    enc = [sp_encoding(np.cumsum(rng.normal(size=L)) * rng.uniform(0.2, 20))[0] for L in lens]
    Lmax = max(e.shape[0] for e in enc)
    x = np.zeros((len(lens), Lmax, 1))
    for i, e in enumerate(enc):
        x[i, : e.shape[0]] = e
    return torch.from_numpy(x).to(torch.float32)


def _tiny_encoder(dtype, hidden=256, mode=1, layers=5, seed=3):
    from chatts_b200.ts_encoder import TimeSeriesEmbedding
    cfg = dict(patch_size=16, num_layers=layers, hidden_size=hidden, num_features=2, max_sequence_length=1024,
               use_position_embedding=mode == 1, use_position_idx=mode == 2, embedding_dim=16)
    in0 = 16 * (1 + 16) if mode == 1 else (32 if mode == 2 else 16)
    g = torch.Generator().manual_seed(seed)
    w = {}
    if mode == 1:
        w["ts_encoder.position_embedding.weight"] = torch.randn(1025, 16, generator=g) * 0.5
    k = in0
    for li in range(layers):
        w[f"ts_encoder.mlp.{2 * li}.weight"] = torch.randn(hidden, k, generator=g) * (1.5 / k ** 0.5)
        w[f"ts_encoder.mlp.{2 * li}.bias"] = torch.randn(hidden, generator=g) * 0.1
        k = hidden
    return cfg, w, TimeSeriesEmbedding(cfg, w, dtype=dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("lens", [[256] * 8, [16], [256, 100, 37, 1000], [256] * 16, [1], [64] * 4 + [48]])
def test_fused_encoder_matches_the_multi_launch_path_and_the_oracle(lens, dtype):
    """Same rows, same rounding points; the K partition of the sums differs (cluster split vs the GEMM's own split), so the two
    paths agree to fp32 summation order.  Row counts 128 (the metric prompt), 1, 89 (ragged), 256 (the upper limit), ..."""
    cfg, w, enc = _tiny_encoder(dtype)
    x = _series_batch(lens)
    assert enc.use_fused
    l0 = enc.ctx.launches
    fused, pc = enc.encode(x.cuda())
    n_fused = enc.ctx.launches - l0
    enc.use_fused = False
    l0 = enc.ctx.launches
    multi, pc2 = enc.encode(x.cuda())
    n_multi = enc.ctx.launches - l0
    torch.cuda.synchronize()
    assert pc.tolist() == pc2.tolist() == [(L + 15) // 16 for L in lens]
    assert n_fused == 3 and n_multi > n_fused            # count + scan + ONE launch for everything else
    wo = {k[len("ts_encoder."):]: v.to(dtype) for k, v in w.items()}
    ref, _ = ote.forward(x.to(dtype), cfg, wo)
    e_paths, e_oracle = rel_err(fused, multi), rel_err(fused, ref)
    record("ts_encoder_fused", rows=int(pc.sum()), dtype=str(dtype), fused_vs_multi_launch=e_paths, fused_vs_oracle=e_oracle,
           multi_launch_vs_oracle=rel_err(multi, ref))
    tol = 6e-3 if dtype == torch.bfloat16 else 8e-4
    assert e_paths < tol and e_oracle < 1.5 * tol and torch.isfinite(fused.float()).all()


@pytest.mark.parametrize("mode", [0, 2])
def test_fused_encoder_other_position_modes(mode):
    cfg, w, enc = _tiny_encoder(torch.bfloat16, mode=mode, layers=3)
    x = _series_batch([256, 64, 128], seed=5)                # multiples of the patch size (the reference raises otherwise, :128)
    fused, pc = enc.encode(x.cuda())
    wo = {k[len("ts_encoder."):]: v.to(torch.bfloat16) for k, v in w.items()}
    ref, _ = ote.forward(x.to(torch.bfloat16), cfg, wo)
    assert pc.tolist() == [16, 4, 8] and rel_err(fused, ref) < 9e-3


def test_fused_encoder_scatters_rows_through_row_map():
    """The last layer writes row i to out[row_map[i]] (the sp-mask scatter into inputs_embeds, chatts_vllm.py:569-573): mapped rows
    hold the features, every other row of the destination is untouched, a negative entry drops its row."""
    cfg, w, enc = _tiny_encoder(torch.bfloat16)
    x = _series_batch([256, 100, 37])
    feats, pc = enc.encode(x.cuda())
    total = int(pc.sum())
    g = torch.Generator().manual_seed(1)
    perm = torch.randperm(total + 9, generator=g)[:total].to(torch.int32)
    perm[3] = -1
    big = torch.full((total + 9, enc.hidden_size), 7.0, device="cuda", dtype=torch.bfloat16)
    counts = enc.patch_counts(x.cuda())
    enc.encode(x.cuda(), out=big, row_map=perm.cuda(), counts=counts)
    torch.cuda.synchronize()
    keep = perm >= 0
    assert torch.equal(big[perm[keep].long().cuda()], feats[keep.cuda()])
    untouched = torch.ones(total + 9, dtype=torch.bool)
    untouched[perm[keep].long()] = False
    assert bool((big[untouched.cuda()] == 7.0).all())


def test_more_than_256_rows_take_the_multi_launch_path():
    cfg, w, enc = _tiny_encoder(torch.bfloat16)
    x = _series_batch([256] * 17)                            # 272 rows
    l0 = enc.ctx.launches
    feats, pc = enc.encode(x.cuda())
    assert feats.shape[0] == 272 and enc.ctx.launches - l0 > 3
