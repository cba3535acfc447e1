"""Pin the oracle against fixtures produced by the reference itself (tests/golden/make_golden.py)
and against the known answer in demo/demo_lora.ipynb cell 6."""
import os

import numpy as np
import pytest
import torch

from oracle import sp_encoding as osp
from oracle import ts_encoder as ote


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_sp_encoding_matches_reference_outputs(golden_dir):
    g = _load(golden_dir, "sp_encoding.npz")
    for i in range(int(g["n"])):
        enc, off, sc = osp.sp_encoding(g[f"in_{i}"])
        assert np.array_equal(enc, g[f"out_{i}"]), f"series {i}"          # bit-exact float64
        assert off == g[f"meta_{i}"][0] and sc == g[f"meta_{i}"][1]
        assert osp.legacy_prefix(off, sc) == str(g[f"prompt_{i}"])


def test_sp_batch_pad_matches_reference(golden_dir):
    g = _load(golden_dir, "sp_encoding.npz")
    encs = [osp.sp_encoding(g[f"in_{i}"])[0] for i in (0, 4, 2)]
    assert np.array_equal(osp.pad_batch(encs), g["batch_out"])


def test_sp_known_answer_demo_notebook():
    """demo/demo_lora.ipynb cell 6: the released processor printed these prefixes."""
    x = np.arange(256)
    ts1 = np.sin(x / 10) * 5.0
    ts1[100:] -= 10.0
    ts2 = x * 0.05
    ts2[103] += 10.0
    _, o1, s1 = osp.sp_encoding(ts1)
    _, o2, s2 = osp.sp_encoding(ts2)
    assert osp.hf_prefix(ts1, o1, s1) == ("[offset=6.0772|scaling=3.6917|length=256|max=4.9979|min=-15.0000|"
                                          "left=0.0000|right=-8.2047]<ts><ts/>")
    assert osp.hf_prefix(ts2, o2, s2) == ("[offset=-6.4141|scaling=2.9120|length=256|max=15.1500|min=0.0000|"
                                          "left=0.0000|right=12.7500]<ts><ts/>")


def _cfg_weights(g):
    cfg = {k[4:]: g[k].item() for k in g.files if k.startswith("cfg.")}
    w = {k[2:]: torch.tensor(g[k]) for k in g.files if k.startswith("w.")}
    return cfg, w


@pytest.mark.parametrize("tag", ["posemb", "posidx", "plain", "posemb_p8"])
def test_ts_encoder_matches_reference_outputs(golden_dir, tag):
    g = _load(golden_dir, f"ts_encoder_{tag}.npz")
    cfg, w = _cfg_weights(g)
    x = torch.tensor(g["x"])
    feats, pc = ote.forward(x, cfg, w)
    assert np.array_equal(pc.numpy(), g["patch_cnt"])                    # bit-exact integers
    assert feats.shape == g["feats"].shape
    np.testing.assert_allclose(feats.numpy(), g["feats"], rtol=0, atol=2e-6)   # fp32, same op order up to cat
    valid, pc2 = ote.patch_count(x, cfg)
    assert np.array_equal(valid.numpy(), g["lengths"])
    assert np.array_equal(pc2.numpy(), g["patch_cnt"])


def test_ts_encoder_ragged_without_posemb_raises_like_reference(golden_dir):
    g = _load(golden_dir, "ts_encoder_plain.npz")
    cfg, w = _cfg_weights(g)
    x = torch.zeros(1, 2 * 17, 1)
    x[0, 0:34:2, 0] = 1.0
    x[0, 1:34:2, 0] = 1.0
    with pytest.raises(AttributeError):
        ote.forward(x, cfg, w)


def test_ts_encoder_all_empty(golden_dir):
    g = _load(golden_dir, "ts_encoder_posemb.npz")
    cfg, w = _cfg_weights(g)
    feats, pc = ote.forward(torch.zeros(3, 64, 1), cfg, w)
    assert feats.shape == (0, cfg["hidden_size"]) and pc.tolist() == [0, 0, 0]
