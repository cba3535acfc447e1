"""Oracle: value-preserving ("sp") normalisation.  TEST INFRASTRUCTURE ONLY.

Follows chatts/utils/encoding_utils.py:23-37 (``sp_encoding``) and :65-86
(``eval_prompt_to_encoding`` -- the batch zero-pad), NetManAIOps/ChatTS @ 09fae34.
The prefix *text* of the released HF processor is not in the reference repo; its format is
taken from the known answer printed in demo/demo_lora.ipynb cell 6.
"""
import numpy as np


def sp_encoding(timeseries):
    """encoding_utils.py:23-37.  float64 in, (float64[2L,1], offset, scale) out."""
    ts = np.asarray(timeseries, dtype=np.float64)
    mean = np.mean(ts)                                   # :25
    scaled = ts - mean                                   # :26
    scale_factor = 1.0                                   # :27
    if np.any(np.abs(scaled) >= 3.0):                    # :28
        scale_factor = np.max(np.abs(scaled)) / 3.0      # :30
        scaled = scaled / scale_factor                   # :31
    # :35  interleave (value, 1.0) -> [2L, 1]
    out = np.stack([scaled, np.ones_like(scaled)], axis=-1).reshape(-1, 1)
    return out, float(-mean), float(scale_factor)


def legacy_prefix(offset, scale_factor):
    """Prefix text emitted by the in-repo function (encoding_utils.py:33)."""
    return f"[Value Offset: {offset:.4f}|Value Scaling: {scale_factor:.4f}]<ts><ts/>"


def hf_prefix(timeseries, offset, scale_factor):
    """Prefix text of the released checkpoint's processor, reconstructed from the known answer
    in demo/demo_lora.ipynb cell 6:
    ``[offset=6.0772|scaling=3.6917|length=256|max=4.9979|min=-15.0000|left=0.0000|right=-8.2047]<ts><ts/>``
    (max/min/left/right are raw-series statistics)."""
    ts = np.asarray(timeseries, dtype=np.float64)
    return (f"[offset={offset:.4f}|scaling={scale_factor:.4f}|length={len(ts)}|"
            f"max={ts.max():.4f}|min={ts.min():.4f}|left={ts[0]:.4f}|right={ts[-1]:.4f}]<ts><ts/>")


def pad_batch(encoded_list):
    """encoding_utils.py:78-84: zero-pad a list of [2L_i, 1] arrays to [N, 2Lmax, 1]
    (zeros => mask 0 in the padded tail)."""
    if len(encoded_list) == 0:
        return np.zeros((0, 0, 1), dtype=np.float64)
    max_len = max(a.shape[0] for a in encoded_list)
    out = np.zeros((len(encoded_list), max_len, 1), dtype=np.float64)
    for i, a in enumerate(encoded_list):
        out[i, : a.shape[0], :] = a
    return out
