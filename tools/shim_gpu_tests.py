"""GPU tests of the kernels that need no tensor core / TMA / cluster, executed on a machine WITHOUT a GPU through the "CUDA on CPU" shim
(tests/cuda_on_cpu): `ctx()` returns a Context whose library is the kernels' own source compiled by g++, `.cuda()` is the identity.
Unlike tools/dryrun_train_gpu_tests.py (which checks the TEST LOGIC against the torch double), this executes the KERNEL SOURCE.
TEST INFRASTRUCTURE ONLY.

    python tools/shim_gpu_tests.py [pytest args]
"""
import functools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytest  # noqa: E402
import torch  # noqa: E402

for name in ("empty", "full", "zeros", "ones", "randn", "tensor", "empty_like", "zeros_like", "arange"):
    orig = getattr(torch, name)

    def mk(orig):
        @functools.wraps(orig)
        def f(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k.pop("device")
            return orig(*a, **k)
        return f
    setattr(torch, name, mk(orig))
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.is_available = lambda: True
torch.cuda.current_device = lambda: 0
torch.cuda.is_current_stream_capturing = lambda: False

from tests.cuda_on_cpu.shim import shim_context  # noqa: E402
import tests.gpu_util as gu  # noqa: E402

_ctx = shim_context()
gu.ctx = lambda: _ctx
gu.record = lambda *a, **k: None

# the cases whose every entry point is part of the shim build
SELECT = {
    "test_gpu_zz_b_sampling.py": None,
    "test_gpu_zz_c_train.py": "adamw_and_clip or lora_pack",
    "test_gpu_train_kernels.py": "swiglu or rmsnorm_bwd or qkv_rope_bwd or ce_loss or cross_entropy or gather or wgrad",
}

if __name__ == "__main__":
    extra = sys.argv[1:]
    rc = 0
    for f, k in SELECT.items():
        args = [os.path.join(ROOT, "tests", f), "-q", "-p", "no:cacheprovider", "--runxfail", "-m", "gpu", "-x"] + (["-k", k] if k and "-k" not in extra else []) + extra
        rc |= int(pytest.main(args))
    sys.exit(rc)
