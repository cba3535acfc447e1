"""Host mirror of the reference's ``chatts/utils/encoding_utils.py``: the same function names, argument
meaning and return shapes, so a caller swaps ``from chatts.utils.encoding_utils import ...`` for
``from chatts_b200.encoding_utils import ...``.  Host-side float64 preprocessing of O(L) values per series (not the GPU hot path);
pinned bit for bit by fixtures produced by running the reference itself (tests/golden/make_golden.py -> encoding_utils.json).

  sp_encoding(ts)                      -> (values interleaved with 1.0, [2L, 1]; "[Value Offset: ..|Value Scaling: ..]<ts><ts/>"; meta)   :23-37
  minmax_scale_encoding(ts)            -> ([L, 1]; "[Offset: ..|Scaled by: ..]<ts><ts/>"; meta)                                         :10-21
  no_encoding(ts)                      -> (array, "<ts><ts/>", {})                                                                      :39-40
  timeseries_encoding(ts, method)      -> dispatch on 'sp' / 'minmax_scale' / 'no'; anything else NotImplementedError                    :42-50
  eval_prompt_to_encoding(p, tss, m)   -> (prompt with each <ts><ts/> prefixed by its numbers, zero-padded batch [N, Lmax, 1])          :65-86
  timeseries_prompt(p, tss)            -> the series written INTO the prompt as rounded nested lists between <ts> and <ts/>               :52-63
  timeseries_to_list(ts, digits, cp)   -> nested python lists of floats rounded to 6 digits                                             :88-103
  extract_and_remove_ts(s)             -> the inverse used by the streaming chat loop (chatts/utils/vllm_stream_qa.py:41-50)
"""
import copy
import json
import re

import numpy as np

from .processor import TS_PLACEHOLDER
from .processor import sp_encoding as _sp_values


def _centre_and_scale(timeseries):
    ts = np.asarray(timeseries)
    mean = np.mean(ts)
    centred = ts - mean
    factor = 1.0
    if np.any(np.abs(centred) >= 3.0):
        factor = np.max(np.abs(centred)) / 3.0
        centred = centred / factor
    return centred, mean, factor


def sp_encoding(timeseries):
    enc, meta = _sp_values(timeseries)
    return enc, f"[Value Offset: {meta['offset']:.4f}|Value Scaling: {meta['scale_factor']:.4f}]{TS_PLACEHOLDER}", meta


def minmax_scale_encoding(timeseries):
    centred, mean, factor = _centre_and_scale(timeseries)
    return (centred[:, np.newaxis], f"[Offset: {-mean:.4f}|Scaled by: {factor:.4f}]{TS_PLACEHOLDER}",
            {"offset": float(-mean), "scale_factor": float(factor)})


def no_encoding(timeseries):
    return np.array(timeseries), TS_PLACEHOLDER, {}


def timeseries_encoding(timeseries, method):
    table = {"minmax_scale": minmax_scale_encoding, "sp": sp_encoding, "no": no_encoding}
    if method not in table:
        raise NotImplementedError(f"Timeseries encoding method: {method} not implemented!")
    return table[method](timeseries)


def _split(prompt, n_series):
    parts = prompt.split(TS_PLACEHOLDER)
    assert n_series == len(parts) - 1            # the reference asserts; callers rely on the AssertionError
    return parts


def timeseries_prompt(prompt, timeseries):
    if isinstance(timeseries, np.ndarray):
        timeseries = timeseries.tolist()
    parts = _split(prompt, len(timeseries))
    out = parts[0]
    for series, tail in zip(timeseries, parts[1:]):
        out += f"<ts>{[[round(v, 3) for v in row] for row in list(series)]}<ts/>" + tail
    return out


def eval_prompt_to_encoding(prompt, timeseries, method):
    parts = _split(prompt, len(timeseries))
    text, encoded = parts[0], []
    for series, tail in zip(timeseries, parts[1:]):
        values, prefix, _ = timeseries_encoding(np.array(series), method)
        text += prefix + tail
        encoded.append(np.array([values]))
    longest = max(a.shape[1] for a in encoded)
    batch = np.concatenate([np.pad(a, ((0, 0), (0, longest - a.shape[1]), (0, 0))) for a in encoded], axis=0)
    return text, batch


def timeseries_to_list(timeseries, digits=6, cp=True):
    data = copy.deepcopy(timeseries) if cp else timeseries
    if isinstance(data, np.ndarray):
        data = data.tolist()
    if type(data[0]) == float:                  # noqa: E721  (the reference tests the exact type: numpy scalars and ints recurse / fail as there)
        for i, v in enumerate(data):
            data[i] = round(float(v), 6)          # the reference rounds to 6 whatever `digits` says
    else:
        for i, v in enumerate(data):
            data[i] = timeseries_to_list(v, digits, cp=False)
    return data


def extract_and_remove_ts(s):
    """'... <ts>[1, 2, 3]<ts/> ...' -> ('... <ts><ts/> ...', [[1, 2, 3]]); None instead of an empty list."""
    found = re.findall(r"(<ts>)(.*?)(<ts/>)", s)
    series = [json.loads(m[1]) for m in found]
    return re.sub(r"(<ts>)(.*?)(<ts/>)", r"\1\3", s), (series or None)
