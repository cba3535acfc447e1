"""Host-side index arithmetic of the ``<ts>`` protocol: where every text token and every TS patch row lands
in the merged embedding sequence.  Pure integer work on the (CPU) input_ids; the device only consumes the
resulting int32 maps (embedding gather ids, patch-row scatter map, positions, cu_seqlens, KV slots).

Layouts (see oracle/merge.py for the slow statement these are checked against, bit-exact):
  * "hf"   -- input_ids keep the un-expanded ``<ts><ts/>`` pair.  Two merge modes, because the checkpoint's remote code
              (``_merge_input_ids_with_time_series_features``) is NOT in the reference repo and cannot be pinned offline:
              mode "insert" (default): both special tokens keep their embeddings and the P_i patch rows
              of the i-th series are INSERTED between them -> P_i + 2 positions per series (README.md:103;
              chatts/utils/inference_tsmllm_deepspeed.py:86,110);
              mode "overwrite": the pair is REPLACED by the P_i patch rows (the two special-token slots are consumed,
              P_i - 2 new positions) -- the LLaVA-style merge, and what the reference's own vLLM path does with the pair
              (chatts_vllm.py:438-444).  ``config.ts_merge_mode`` / ``CTS_TS_MERGE_MODE`` selects it.
  * "vllm" -- the prompt already holds P_i copies of ``<ts>``; those positions are OVERWRITTEN in order
              (chatts/vllm/chatts_vllm.py:405-415,569-573).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class MergedLayout:
    ids: np.ndarray          # int32 [T]  token id per merged position, -1 where a patch row goes
    row_map: np.ndarray      # int32 [sum P]  merged position of every patch row (series order, then patch order)
    positions: np.ndarray    # int32 [T]  rotary position of every merged position (0.. per sample)
    cu_seqlens: np.ndarray   # int32 [B+1]
    src_col: np.ndarray      # int32 [T]  column of input_ids the position came from (-1 for patch rows)

    @property
    def total(self):
        return int(self.ids.shape[0])

    @property
    def lens(self):
        return np.diff(self.cu_seqlens)


def hf_layout(input_ids, attention_mask, patch_cnt, ts_start, mode="insert"):
    if mode not in ("insert", "overwrite"):
        raise ValueError(f"ts merge mode must be 'insert' or 'overwrite', got {mode!r}")
    ids = np.asarray(input_ids)
    if ids.ndim == 1:
        ids = ids[None]
    am = np.ones_like(ids) if attention_mask is None else np.asarray(attention_mask)
    if am.ndim == 1:
        am = am[None]
    patch_cnt = np.asarray(patch_cnt, dtype=np.int64).reshape(-1)
    B, S = ids.shape
    keep = am.astype(bool)
    flat_ids = ids[keep]                                   # batch-major, prompt order
    flat_col = np.broadcast_to(np.arange(S), (B, S))[keep]
    n_real = keep.sum(axis=1)
    is_ts = flat_ids == ts_start
    n_ts = int(is_ts.sum())
    if n_ts != patch_cnt.shape[0]:
        # the reference asserts the same thing (chatts/utils/encoding_utils.py:58,68)
        raise AssertionError(f"{n_ts} <ts> placeholders in the batch but {patch_cnt.shape[0]} time series were given")
    if mode == "insert":
        extra = np.zeros(flat_ids.shape[0], dtype=np.int64)
        extra[is_ts] = patch_cnt
        span = 1 + extra
        is_text = np.ones(flat_ids.shape[0], dtype=bool)
        first = 1                                           # the rows start right after the kept <ts> token
    else:
        # "overwrite": <ts> and <ts/> give up their slots; the i-th <ts> position expands to P_i rows, <ts/> to nothing
        is_text = ~(is_ts | (flat_ids == ts_start + 1))
        span = is_text.astype(np.int64)
        span[is_ts] = patch_cnt
        first = 0
    tok_pos = np.cumsum(span) - span                       # merged flat position of every input token (start of its span)
    T = int(span.sum())
    out_ids = np.full(T, -1, dtype=np.int32)
    out_ids[tok_pos[is_text]] = flat_ids[is_text].astype(np.int32)
    src_col = np.full(T, -1, dtype=np.int32)
    src_col[tok_pos[is_text]] = flat_col[is_text].astype(np.int32)
    # patch rows: global row r of series k sits at tok_pos[<ts>_k] + first + (r - first_row_k)
    first_row = np.cumsum(patch_cnt) - patch_cnt
    total_rows = int(patch_cnt.sum())
    series_of_row = np.repeat(np.arange(n_ts), patch_cnt)
    within = np.arange(total_rows) - first_row[series_of_row]
    row_map = (tok_pos[is_ts][series_of_row] + first + within).astype(np.int32)
    # per-sample lengths
    sample_of_tok = np.repeat(np.arange(B), n_real)
    lens = np.bincount(sample_of_tok, weights=span, minlength=B).astype(np.int64)
    cu = np.zeros(B + 1, dtype=np.int32)
    cu[1:] = np.cumsum(lens)
    positions = (np.arange(T) - np.repeat(cu[:-1].astype(np.int64), lens)).astype(np.int32)
    return MergedLayout(out_ids, row_map, positions, cu, src_col)


def vllm_layout(token_ids, total_rows, ts_start, cu_seqlens=None):
    ids = np.asarray(token_ids).reshape(-1)
    pos = np.nonzero(ids == ts_start)[0]
    if pos.shape[0] != total_rows:
        raise ValueError(f"{pos.shape[0]} placeholder positions but {total_rows} patch rows")
    out_ids = ids.astype(np.int32).copy()
    out_ids[pos] = -1
    cu = np.array([0, ids.shape[0]], dtype=np.int32) if cu_seqlens is None else np.asarray(cu_seqlens, dtype=np.int32)
    lens = np.diff(cu)
    positions = (np.arange(ids.shape[0]) - np.repeat(cu[:-1].astype(np.int64), lens)).astype(np.int32)
    src = np.arange(ids.shape[0], dtype=np.int32)
    src[pos] = -1
    return MergedLayout(out_ids, pos.astype(np.int32), positions, cu, src)


def expand_prompt_vllm(token_ids, series_token_lists, patch_cnt, ts_start):
    """chatts_vllm.py:405-415,438-444: each [<ts>, <ts/>] pair -> ts_tokens_i padded with <ts> up to patch_cnt_i."""
    out, i, k = [], 0, 0
    toks = list(token_ids)
    while i < len(toks):
        if i + 1 < len(toks) and toks[i] == ts_start and toks[i + 1] == ts_start + 1:
            rep = list(series_token_lists[k])
            have = rep.count(ts_start)
            if have < patch_cnt[k]:
                rep += [ts_start] * (int(patch_cnt[k]) - have)
            out += rep
            k += 1
            i += 2
        else:
            out.append(toks[i])
            i += 1
    return out
