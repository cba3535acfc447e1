import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **kv):
    """Append a measured parity number to gpurun_out/parity_metrics.jsonl (travels back from the GPU box)."""
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_metrics.jsonl"), "a") as f:
        f.write(json.dumps(dict(name=name, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in kv.items()})) + "\n")


def rel_err(a, ref):
    """max|a - ref| / max|ref|  (the north_star's logit metric)."""
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def ctx():
    from chatts_b200 import _cabi
    return _cabi.get_context()
