"""CPU: chatts_b200.encoding_utils (host mirror of the reference's chatts/utils/encoding_utils.py, rows A1 / A2) against outputs of
the reference itself (tests/golden/encoding_utils.json, produced by tests/golden/make_golden.py): every encoding method, ragged
batches through the zero pad, the prompt/text helpers.  Bit-exact -- float64 values compared with ==, strings with ==."""
import json
import os

import numpy as np
import pytest

from chatts_b200 import encoding_utils as eu


@pytest.fixture(scope="module")
def cases(golden_dir):
    return json.load(open(os.path.join(golden_dir, "encoding_utils.json")))


def test_eval_prompt_to_encoding_every_method(cases):
    n = 0
    for c in cases:
        if c["fn"] != "eval_prompt_to_encoding":
            continue
        text, batch = eu.eval_prompt_to_encoding(c["prompt"], c["timeseries"], c["method"])
        assert text == c["out_prompt"]
        assert list(batch.shape) == c["out_shape"]
        assert np.array_equal(batch, np.array(c["out_batch"], dtype=np.float64).reshape(c["out_shape"]))
        n += 1
    assert n == 12


def test_timeseries_encoding_and_metadata(cases):
    for c in cases:
        if c["fn"] != "timeseries_encoding":
            continue
        enc, prefix, meta = eu.timeseries_encoding(np.array(c["timeseries"]), c["method"])
        assert prefix == c["out_prompt"] and meta == c["meta"]
        assert np.array_equal(enc, np.array(c["out"], dtype=np.float64))
    with pytest.raises(NotImplementedError):
        eu.timeseries_encoding(np.zeros(4), "zscore")
    enc, prefix, meta = eu.no_encoding([1.0, 2.0])
    assert prefix == "<ts><ts/>" and meta == {} and enc.tolist() == [1.0, 2.0]


def test_text_helpers(cases):
    for c in cases:
        if c["fn"] == "timeseries_prompt":
            ts = np.array(c["timeseries"]) if c.get("as_array") else c["timeseries"]
            assert eu.timeseries_prompt(c["prompt"], ts) == c["out"]
        elif c["fn"] == "timeseries_to_list":
            obj = np.array(c["in"]) if c.get("as_array") else c["in"]
            assert eu.timeseries_to_list(obj) == c["out"]
    with pytest.raises(AssertionError):                      # placeholder / series count mismatch (encoding_utils.py:58,68)
        eu.eval_prompt_to_encoding("only one <ts><ts/>", [[1.0], [2.0]], "sp")
    with pytest.raises(AssertionError):
        eu.timeseries_prompt("none", [[[1.0]]])
    src = [1.0, 2.0]
    out = eu.timeseries_to_list(src)
    assert out == src and out is not src                     # cp=True works on a copy


def test_extract_and_remove_ts_round_trip():
    s = 'compare <ts>[1, 2.5, 3]<ts/> with <ts>[[0.5], [1.5]]<ts/> please'
    text, series = eu.extract_and_remove_ts(s)
    assert text == "compare <ts><ts/> with <ts><ts/> please" and series == [[1, 2.5, 3], [[0.5], [1.5]]]
    assert eu.extract_and_remove_ts("no series here") == ("no series here", None)
