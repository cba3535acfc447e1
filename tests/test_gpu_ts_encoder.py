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
