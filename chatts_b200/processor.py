"""ChatTSProcessor -- host mirror of the checkpoint's AutoProcessor (remote code, NOT in the reference repo;
SURVEY.md §8f N1).  What is pinned: the value-preserving normalisation (chatts/utils/encoding_utils.py:23-37),
the prefix numbers/format (demo/demo_lora.ipynb cell 6), the flat ``timeseries`` list consumed in ``<ts><ts/>``
order across the batch (chatts/utils/inference_tsmllm_deepspeed.py:75-89), zero padding of the series tensor
(encoding_utils.py:78-84), and the call signature ``processor(text=[...], timeseries=[...], padding=True,
return_tensors="pt")`` (README.md:98).  Tokenisation is delegated to the tokenizer handed in, exactly like
``AutoProcessor.from_pretrained(path, tokenizer=tokenizer)`` (README.md:90).

This is host-side float64 preprocessing on O(L) values per series -- not the GPU hot path.
"""
import numpy as np
import torch

TS_PLACEHOLDER = "<ts><ts/>"


def sp_encoding(timeseries):
    """encoding_utils.py:23-37: mean-centre; scale so max|x| = 3 if any |x| >= 3; interleave (value, 1.0)."""
    ts = np.asarray(timeseries, dtype=np.float64)
    mean = ts.mean()
    scaled = ts - mean
    scale = 1.0
    if np.any(np.abs(scaled) >= 3.0):
        scale = np.abs(scaled).max() / 3.0
        scaled = scaled / scale
    enc = np.stack([scaled, np.ones_like(scaled)], axis=-1).reshape(-1, 1)
    return enc, {"offset": float(-mean), "scale_factor": float(scale)}


def render_prefix(ts, meta):
    """demo/demo_lora.ipynb cell 6: ``[offset=6.0772|scaling=3.6917|length=256|max=4.9979|min=-15.0000|left=0.0000|right=-8.2047]<ts><ts/>``."""
    ts = np.asarray(ts, dtype=np.float64)
    return (f"[offset={meta['offset']:.4f}|scaling={meta['scale_factor']:.4f}|length={len(ts)}|max={ts.max():.4f}|"
            f"min={ts.min():.4f}|left={ts[0]:.4f}|right={ts[-1]:.4f}]{TS_PLACEHOLDER}")


class SimpleTokenizer:
    """Deterministic stand-in used when no checkpoint tokenizer exists offline: UTF-8 bytes -> ids [0,256),
    ``<ts>`` / ``<ts/>`` -> the config's special ids.  Same call surface as an HF tokenizer for what the
    processor needs (left padding, attention mask, decode)."""
    padding_side = "left"

    def __init__(self, ts_start, pad_token_id, eos_token_id=None):
        self.ts_start, self.pad_token_id, self.eos_token_id = ts_start, pad_token_id, eos_token_id

    def encode(self, text):
        out, i = [], 0
        while i < len(text):
            if text.startswith("<ts/>", i):
                out.append(self.ts_start + 1)
                i += 5
            elif text.startswith("<ts>", i):
                out.append(self.ts_start)
                i += 4
            else:
                out.extend(text[i].encode("utf-8"))
                i += 1
        return out

    def __call__(self, text, padding=True, return_tensors="pt", **_):
        seqs = [self.encode(t) for t in ([text] if isinstance(text, str) else text)]
        n = max(len(s) for s in seqs)
        ids = np.full((len(seqs), n), self.pad_token_id, dtype=np.int64)
        am = np.zeros((len(seqs), n), dtype=np.int64)
        for i, s in enumerate(seqs):
            if self.padding_side == "left":
                ids[i, n - len(s):], am[i, n - len(s):] = s, 1
            else:
                ids[i, : len(s)], am[i, : len(s)] = s, 1
        return {"input_ids": torch.from_numpy(ids), "attention_mask": torch.from_numpy(am)}

    def decode(self, ids, skip_special_tokens=True):
        b = bytes(int(t) for t in ids if 0 <= int(t) < 256)
        return b.decode("utf-8", errors="replace")


class ChatTSProcessor:
    def __init__(self, tokenizer, config=None, dtype=torch.float32):
        self.tokenizer, self.config, self.dtype = tokenizer, config, dtype

    @classmethod
    def from_pretrained(cls, path, tokenizer=None, trust_remote_code=True, **kw):
        from .config import ChatTSConfig
        return cls(tokenizer, ChatTSConfig.from_json(path), **kw)

    def encode_series(self, series):
        encs, prefixes = [], []
        for ts in series:
            if not isinstance(ts, (np.ndarray, list, tuple, torch.Tensor)):
                raise TypeError(f"Unsupported time series type: {type(ts)}")       # chatts_vllm.py:277-279
            ts = np.asarray(ts, dtype=np.float64)
            if ts.ndim != 1:
                raise ValueError("each time series must be one-dimensional")
            enc, meta = sp_encoding(ts)
            encs.append(enc)
            prefixes.append(render_prefix(ts, meta))
        return encs, prefixes

    @staticmethod
    def render_text(text, prefixes):
        """Replace the i-th ``<ts><ts/>`` of ``text`` by the i-th rendered prefix (which ends in ``<ts><ts/>`` itself)."""
        parts = text.split(TS_PLACEHOLDER)
        assert len(parts) - 1 == len(prefixes), "time series / <ts><ts/> placeholder count mismatch"
        s = parts[0]
        for j, pre in enumerate(prefixes):
            s += pre + parts[j + 1]
        return s

    def pad_series(self, encs):
        """encoding_utils.py:78-84: zero-padded [N, 2 Lmax, 1] tensor of the encoded series."""
        if not encs:
            return torch.zeros(0, 0, 1, dtype=self.dtype)
        max_len = max(e.shape[0] for e in encs)
        arr = np.zeros((len(encs), max_len, 1), dtype=np.float64)
        for i, e in enumerate(encs):
            arr[i, : e.shape[0]] = e
        return torch.from_numpy(arr).to(self.dtype)

    def __call__(self, text, timeseries=None, padding=True, return_tensors="pt", vllm_flag=False, **kw):
        if isinstance(text, str):
            text = [text]
        timeseries = [] if timeseries is None else list(timeseries)
        encs, prefixes = self.encode_series(timeseries)
        k, rendered = 0, []
        for t in text:
            n = t.count(TS_PLACEHOLDER)
            assert k + n <= len(prefixes), "more <ts><ts/> placeholders than time series"    # encoding_utils.py:58,68
            rendered.append(self.render_text(t, prefixes[k: k + n]))
            k += n
        assert k == len(prefixes), "time series / <ts><ts/> placeholder count mismatch"
        if vllm_flag:
            # chatts_vllm.py:319-348,392: per series (ts_tokens, encoded [1, 2L, 1])
            toks = [self.tokenizer.encode(p) if hasattr(self.tokenizer, "encode") else None for p in prefixes]
            return {"timeseries": [(tk, e[None]) for tk, e in zip(toks, encs)], "text": rendered}
        out = dict(self.tokenizer(rendered, padding=padding, return_tensors=return_tensors))
        out["timeseries"] = self.pad_series(encs)                                   # encoding_utils.py:78-84
        return out
