"""Oracle: splice TS patch rows into the token stream.  TEST INFRASTRUCTURE ONLY.

Two layouts exist for the same protocol (``<ts>`` = ts_token_start_index, ``<ts/>`` = +1,
chatts/vllm/chatts_vllm.py:441):

* ``vllm``  (chatts_vllm.py:405-415, 564-574): the processor has already expanded each ``<ts><ts/>`` pair
  into P copies of ``<ts>``; every position whose id == ``<ts>`` is OVERWRITTEN, in order, by the
  concatenated patch rows (``merge_multimodal_embeddings``).
* ``hf``    (remote code of the checkpoint, NOT in the reference repo -- PARITY UNPINNED): ``input_ids``
  keep the un-expanded ``<ts><ts/>`` pair (README.md:103: generate() returns the original ids;
  demo/demo_lora.ipynb cell 6 output shows the pair adjacent in the echoed prompt) and the model
  INSERTS the P_i patch rows of the i-th series between them, so the merged length is
  S_text + sum(P) (token accounting in chatts/utils/inference_tsmllm_deepspeed.py:86,110).  Series are
  consumed in prompt order across the flattened batch (inference_tsmllm_deepspeed.py:75-80).
  ``mode="overwrite"`` states the other candidate for the absent remote code: the ``<ts><ts/>`` pair gives up its two
  slots and the P_i rows take their place (P_i - 2 new positions; every ``<ts>`` / ``<ts/>`` token is dropped), which is
  also what the reference's vLLM path does with the pair (chatts_vllm.py:438-444).  Which of the two the released
  checkpoints use can only be settled against a real checkpoint -- both are kept, ``insert`` is the default.

Pure-Python loops on purpose: this is the slow, obviously-right statement the vectorised host code in
chatts_b200/layout.py is checked against (bit-exact indices).
"""
import torch


def hf_layout(input_ids, attention_mask, patch_cnt, ts_start, mode="insert"):
    """input_ids/attention_mask: [B,S] (any padding side); patch_cnt: list[int], one per ``<ts>`` in
    batch-major prompt order.  Returns per sample a list of entries, one per merged position:
        ("tok", column_in_input_ids)  or  ("ts", global_patch_row)
    with padded (mask==0) columns dropped."""
    ids = torch.as_tensor(input_ids).tolist()
    am = torch.as_tensor(attention_mask).tolist()
    series, row = 0, 0
    out = []
    for b in range(len(ids)):
        ent = []
        for s in range(len(ids[b])):
            if not am[b][s]:
                continue
            if mode == "insert" or ids[b][s] not in (ts_start, ts_start + 1):
                ent.append(("tok", s))
            if ids[b][s] == ts_start:
                assert series < len(patch_cnt), "more <ts> tokens than series"   # encoding_utils.py:58,68
                for _ in range(int(patch_cnt[series])):
                    ent.append(("ts", row))
                    row += 1
                series += 1
        out.append(ent)
    assert series == len(patch_cnt), "series / <ts> count mismatch"
    return out


def hf_merge(input_ids, attention_mask, embed_table, ts_feats, patch_cnt, ts_start, mode="insert"):
    """Per-sample merged embeddings [T_b, H] in the HF layout (``mode``: insert | overwrite)."""
    lay = hf_layout(input_ids, attention_mask, patch_cnt, ts_start, mode)
    ids = torch.as_tensor(input_ids)
    res = []
    for b, ent in enumerate(lay):
        rows = []
        for kind, idx in ent:
            rows.append(embed_table[ids[b, idx]] if kind == "tok" else ts_feats[idx].to(embed_table.dtype))
        res.append(torch.stack(rows) if rows else embed_table.new_zeros(0, embed_table.shape[1]))
    return res


def vllm_merge(input_ids, embed_table, ts_feats, ts_start):
    """chatts_vllm.py:569-573: flat token stream [T]; rows at ids == <ts> are overwritten in order."""
    ids = torch.as_tensor(input_ids).reshape(-1)
    emb = embed_table[ids].clone()
    pos = (ids == ts_start).nonzero().reshape(-1)
    assert pos.numel() == ts_feats.shape[0], "placeholder / patch-row count mismatch"
    emb[pos] = ts_feats.to(emb.dtype)
    return emb


def vllm_expand_prompt(token_ids, series_token_lists, patch_cnt, ts_start):
    """chatts_vllm.py:405-415,438-444: replace the i-th [<ts>, <ts/>] pair by ts_tokens_i extended with
    ``<ts>`` until it holds patch_cnt_i placeholders."""
    out, i, k = [], 0, 0
    toks = list(token_ids)
    while i < len(toks):
        if i + 1 < len(toks) and toks[i] == ts_start and toks[i + 1] == ts_start + 1:
            rep = list(series_token_lists[k])
            have = sum(1 for t in rep if t == ts_start)
            if have < patch_cnt[k]:
                rep.extend([ts_start] * (patch_cnt[k] - have))
            out.extend(rep)
            k += 1
            i += 2
        else:
            out.append(toks[i])
            i += 1
    return out
