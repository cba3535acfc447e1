"""Processor (host logic) against the reference-generated fixtures and the notebook known answer."""
import os

import numpy as np
import pytest
import torch

from chatts_b200 import ChatTSConfig, ChatTSProcessor, SimpleTokenizer, sp_encoding
from chatts_b200.processor import render_prefix


def test_sp_encoding_bit_exact_vs_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "sp_encoding.npz"))
    for i in range(int(g["n"])):
        enc, meta = sp_encoding(g[f"in_{i}"])
        assert np.array_equal(enc, g[f"out_{i}"])
        assert meta["offset"] == g[f"meta_{i}"][0] and meta["scale_factor"] == g[f"meta_{i}"][1]


def test_prefix_known_answer():
    x = np.arange(256)
    ts1 = np.sin(x / 10) * 5.0
    ts1[100:] -= 10.0
    enc, meta = sp_encoding(ts1)
    assert render_prefix(ts1, meta) == ("[offset=6.0772|scaling=3.6917|length=256|max=4.9979|min=-15.0000|left=0.0000|"
                                        "right=-8.2047]<ts><ts/>")


def test_call_contract(golden_dir):
    cfg = ChatTSConfig.tiny()
    tok = SimpleTokenizer(cfg.ts_token_start_index, cfg.pad_token_id)
    proc = ChatTSProcessor(tok, cfg)
    a, b = np.linspace(0, 1, 40), np.linspace(5, -5, 100)
    out = proc(text=["p <ts><ts/> q <ts><ts/>", "r"], timeseries=[a, b], padding=True, return_tensors="pt")
    assert set(out) >= {"input_ids", "attention_mask", "timeseries"}
    assert out["timeseries"].shape == (2, 200, 1)
    ids, am = out["input_ids"], out["attention_mask"]
    assert ids.shape == am.shape and am[1, 0] == 0 and am[1, -1] == 1          # left padding
    assert int((ids == cfg.ts_token_start_index).sum()) == 2 and int((ids == cfg.ts_token_start_index + 1).sum()) == 2
    # zero padding => mask 0 in the tail of the shorter series (encoding_utils.py:78-84)
    assert out["timeseries"][0, 80:, 0].abs().sum() == 0 and out["timeseries"][0, 1:80:2, 0].min() == 1
    with pytest.raises(AssertionError):
        proc(text=["<ts><ts/>"], timeseries=[a, b])
    with pytest.raises(TypeError):
        proc(text=["<ts><ts/>"], timeseries=["nope"])
    v = proc(text=["<ts><ts/>"], timeseries=[a], vllm_flag=True)
    assert v["timeseries"][0][1].shape == (1, 80, 1)
