"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only where /root/reference exists (the build container, never the GPU box):

    python tests/golden/make_golden.py

* ``TimeSeriesEmbedding`` and ``get_patch_cnt`` are AST-extracted from
  /root/reference/chatts/vllm/chatts_vllm.py (the module itself cannot be imported: it targets
  vllm 0.8.5 and fails on the installed 0.22) and EXECUTED on CPU with seeded weights/inputs.
* ``sp_encoding`` / ``eval_prompt_to_encoding`` are imported from chatts.utils.encoding_utils.

Outputs (small files, committed):  ts_encoder_{posemb,posidx,plain}.npz, sp_encoding.npz, encoding_utils.json.
Nothing from the reference's sources is copied into the repo -- only its numeric outputs.
"""
import ast
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_encoder():
    src = open(os.path.join(REF, "chatts/vllm/chatts_vllm.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body
            if (isinstance(n, ast.ClassDef) and n.name == "TimeSeriesEmbedding")
            or (isinstance(n, ast.FunctionDef) and n.name == "get_patch_cnt")]
    assert len(keep) == 2
    mod = ast.Module(body=keep, type_ignores=[])
    ns = {"torch": torch, "nn": torch.nn, "PretrainedConfig": dict, "math": __import__("math")}
    exec(compile(mod, "chatts_vllm_extract", "exec"), ns)
    return ns["TimeSeriesEmbedding"], ns["get_patch_cnt"]


def demo_series():
    """README.md:92-93 / demo/demo_lora.ipynb cell 5 / demo/demo_vllm.py:33-45."""
    x = np.arange(256)
    ts1 = np.sin(x / 10) * 5.0
    ts1[100:] -= 10.0
    ts2 = x * 0.05
    ts2[103] += 10.0
    x3 = np.arange(1000)
    ts3 = x3 * 0.01
    ts3[500] += 8.0
    return [ts1, ts2, ts3]


def random_series(rng, length):
    t = np.arange(length)
    a = rng.uniform(0.5, 50)
    s = a * np.sin(2 * np.pi * t / rng.uniform(16, 128)) + rng.uniform(-0.05, 0.05) * t
    s += rng.normal(0, 0.1 * a, size=length)
    if length >= 8:
        s[rng.integers(length // 4, max(length // 4 + 1, 3 * length // 4)):] += rng.choice([-2, 2]) * a
    return s


def main():
    sys.path.insert(0, REF)
    from chatts.utils import encoding_utils as eu

    TimeSeriesEmbedding, get_patch_cnt = load_reference_encoder()

    # ---------------- sp_encoding ----------------
    rng = np.random.default_rng(7)
    series = demo_series() + [random_series(rng, n) for n in (1, 5, 16, 17, 64, 255, 256, 1024)]
    series += [np.full(32, 2.5), np.linspace(-2.9, 2.9, 40), np.linspace(-3.0, 3.0, 40)]
    rec = {}
    for i, s in enumerate(series):
        enc, prompt, meta = eu.sp_encoding(np.array(s, dtype=np.float64))
        rec[f"in_{i}"] = np.array(s, dtype=np.float64)
        rec[f"out_{i}"] = enc
        rec[f"meta_{i}"] = np.array([meta["offset"], meta["scale_factor"]])
        rec[f"prompt_{i}"] = np.array(prompt)
    # the batch pad of eval_prompt_to_encoding
    prompt = "A <ts><ts/> B <ts><ts/> C <ts><ts/>."
    rp, batch = eu.eval_prompt_to_encoding(prompt, [series[0].tolist(), series[4].tolist(), series[2].tolist()], "sp")
    rec["batch_prompt"] = np.array(rp)
    rec["batch_out"] = batch
    rec["n"] = np.array(len(series))
    np.savez_compressed(os.path.join(HERE, "sp_encoding.npz"), **rec)

    # ---------------- the rest of encoding_utils (every method, ragged batches, the text helpers) ----------------
    import json
    cases = []
    pool = [series[0], series[4], series[2], series[11], series[12], series[13]]
    for method in ("sp", "minmax_scale", "no"):
        for pick in ([0], [1, 2], [3, 4, 5], [0, 1, 2, 3]):
            tss = [pool[i].tolist() for i in pick]
            prompt = "Q" + "".join(f" s{j}: <ts><ts/>" for j in range(len(pick))) + " end."
            if method == "no" and len({len(t) for t in tss}) > 1:
                tss = [t[:20] for t in tss]              # 'no' keeps 1-D series: np.pad over 3 axes needs [1, L, 1]-shaped input
                tss = [[[v] for v in t] for t in tss]
            elif method == "no":
                tss = [[[v] for v in t] for t in tss]
            rp, batch = eu.eval_prompt_to_encoding(prompt, tss, method)
            cases.append({"fn": "eval_prompt_to_encoding", "method": method, "prompt": prompt, "timeseries": tss, "out_prompt": rp,
                          "out_batch": batch.tolist(), "out_shape": list(batch.shape)})
    for s_ in (series[0], series[13]):
        for method in ("sp", "minmax_scale"):
            enc, pr, meta = eu.timeseries_encoding(np.array(s_), method)
            cases.append({"fn": "timeseries_encoding", "method": method, "timeseries": s_.tolist(), "out": enc.tolist(), "out_prompt": pr, "meta": meta})
    two = [[[1.23456, 2.0], [3.14159, -4.5]], [[0.0005, 1e-7], [7.0, 8.12345678]]]
    cases.append({"fn": "timeseries_prompt", "prompt": "a <ts><ts/> b <ts><ts/> c", "timeseries": two,
                  "out": eu.timeseries_prompt("a <ts><ts/> b <ts><ts/> c", two)})
    cases.append({"fn": "timeseries_prompt", "prompt": "a <ts><ts/> b <ts><ts/> c", "timeseries": two, "as_array": True,
                  "out": eu.timeseries_prompt("a <ts><ts/> b <ts><ts/> c", np.array(two))})
    for obj in ([0.123456789, 2.5, -1e-9], [[1.00000049, 2.0], [3.3333333333, 4.0]], np.array([[0.1234567891, 5.0]]).tolist()):
        cases.append({"fn": "timeseries_to_list", "in": obj, "out": eu.timeseries_to_list(obj)})
    cases.append({"fn": "timeseries_to_list", "in": [[0.1234567891, 5.0]], "as_array": True, "out": eu.timeseries_to_list(np.array([[0.1234567891, 5.0]]))})
    json.dump(cases, open(os.path.join(HERE, "encoding_utils.json"), "w"))

    # ---------------- TS encoder, three position modes ----------------
    def run(tag, cfg, lengths, seed):
        torch.manual_seed(seed)
        enc = TimeSeriesEmbedding(cfg).eval()
        # nn.Embedding/Linear default inits are fine; make them a little larger so outputs are not tiny
        with torch.no_grad():
            for p in enc.parameters():
                p.mul_(2.0)
        g = np.random.default_rng(seed)
        encs = [eu.sp_encoding(random_series(g, n))[0][None] if n > 0 else np.zeros((1, 0, 1)) for n in lengths]
        max_len = max(max(a.shape[1] for a in encs), 2)
        x = np.zeros((len(encs), max_len, 1), dtype=np.float32)
        for i, a in enumerate(encs):
            x[i, : a.shape[1]] = a
        xt = torch.tensor(x)
        with torch.no_grad():
            feats, pc = enc(xt)
            pc2 = get_patch_cnt(xt, cfg)
        assert torch.equal(pc, pc2)
        out = {"x": x, "feats": feats.numpy(), "patch_cnt": pc.numpy(), "lengths": np.array(lengths)}
        for k, v in enc.state_dict().items():
            out["w." + k] = v.numpy()
        for k, v in cfg.items():
            out["cfg." + k] = np.array(v)
        np.savez_compressed(os.path.join(HERE, f"ts_encoder_{tag}.npz"), **out)
        print(tag, "rows", feats.shape, "patch_cnt", pc.tolist())

    base = dict(patch_size=16, num_layers=3, hidden_size=128, num_features=2, max_sequence_length=512)
    run("posemb", dict(base, use_position_embedding=True, embedding_dim=16),
        [256, 1, 16, 17, 0, 64, 100, 512, 31], seed=11)
    # without pos-emb the reference only accepts multiples of patch_size (AttributeError otherwise)
    run("posidx", dict(base, use_position_idx=True), [256, 16, 0, 64, 512, 32], seed=12)
    run("plain", dict(base), [256, 16, 0, 64, 512, 32], seed=13)
    # a second pos-emb case with different patch size / layer count / embedding dim
    run("posemb_p8", dict(patch_size=8, num_layers=2, hidden_size=64, num_features=2, max_sequence_length=128,
                          use_position_embedding=True, embedding_dim=4), [5, 8, 9, 128, 0, 77], seed=14)


if __name__ == "__main__":
    main()
