"""CPU: kernel SOURCE executed without a GPU.  tests/cuda_on_cpu compiles a few files of chatts_b200/csrc with g++ against a small
"CUDA on CPU" shim (CUDA threads = fibers, __syncthreads / shuffles / atomics provided, blocks one after the other) and the
product's own ctypes wrappers drive them on host memory:

  * tools/shim_gpu_tests.py runs the GPU test cases of the sampling kernel, AdamW + clip, adapter packing, the attention backward
    (attention_bwd.cu: wmma through a fragment shim) and -- in a HYBRID context where the tcgen05 GEMMs and the attention forward are
    answered by the torch double -- the whole LoRA training step against the oracle (all still pending on a B200), plus, as
    calibration of the shim itself, the elementwise / cross-entropy / weight-gradient tests that already passed on a B200
    (the CPU suite runs the --quick subset; the full selection is 56 cases in about 4 minutes);
  * the low-latency all-reduce (allreduce_ll.cu) runs with its ranks as PROCESSES sharing the symmetric regions: real concurrency
    between ranks, torn 16-byte units, consecutive calls without any barrier in between.

Only kernels without tensor cores / TMA / clusters can run this way.  Not a memory-model check (x86 is stronger than PTX)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pending_and_validated_kernels_pass_their_gpu_tests_on_the_shim():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "shim_gpu_tests.py"), "--quick"], capture_output=True, text=True, timeout=850,
                       cwd=ROOT)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    import re
    passed = [int(n) for n in re.findall(r"(\d+) passed", r.stdout)]
    assert len(passed) == 3 and min(passed) >= 7 and "failed" not in r.stdout, tail      # sampling, training (incl. one whole step), validated kernels
    native = r.stdout.split("entry points running from kernel source:")[1].split("\n")[0].split()
    assert {"sample_advance", "attn_bwd", "adamw", "grad_norm_clip", "lora_pack", "lora_wgrad", "ce_loss_grad", "rmsnorm_bwd", "swiglu_bwd",
            "qkv_rope_bwd"} <= set(native), native


def _expected(parts, split, resid, dt):
    acc = torch.zeros_like(parts[0][0])
    for p in parts:
        loc = p[0].clone()
        for s in range(1, split):
            loc = loc + p[s]
        acc = acc + loc
    return (resid.float() + acc.to(dt).float()).to(dt)


@pytest.mark.parametrize("W,T,h,split,dt", [(2, 3, 256, 2, torch.bfloat16), (4, 2, 512, 1, torch.float16), (2, 1, 128, 3, torch.bfloat16)])
def test_ll_allreduce_kernel_with_ranks_as_processes(W, T, h, split, dt):
    from tests.cuda_on_cpu.shim import shim_context
    c = shim_context()
    tmax, calls = 4, [0, 1, 0, 0, 1]                         # buffer set per call (one repeated set: allowed, see the protocol test)
    nbytes = c.peer_ll_region_bytes(W, tmax, h)
    g = torch.Generator().manual_seed(W * 100 + h)
    share = lambda t: t.share_memory_()
    regions = [[share(torch.zeros((nbytes + 3) // 4, dtype=torch.int32)) for _ in range(W)] for _ in range(2)]
    ptrs = [share(torch.tensor([r.data_ptr() for r in regions[b]], dtype=torch.int64)) for b in range(2)]
    state = [share(torch.zeros(2, dtype=torch.int32)) for _ in range(W)]
    h0 = (torch.randn(T, h, generator=g) * 0.5).to(dt)
    nw = (1.0 + 0.1 * torch.randn(h, generator=g)).to(dt)
    parts = [[share((torch.randn(split, T, h, generator=g) * 0.3)) for _ in range(W)] for _ in calls]
    # every rank keeps its own h (updated in place by every call) and logs h / norm_out of every call for the parent to check
    resid = [share(h0.clone()) for _ in range(W)]
    xn = [share(torch.zeros(T, h, dtype=dt)) for _ in range(W)]
    log_h = [share(torch.zeros(len(calls), T, h, dtype=dt)) for _ in range(W)]
    log_n = [share(torch.zeros(len(calls), T, h, dtype=dt)) for _ in range(W)]
    pids = []
    for r in range(W):
        pid = os.fork()
        if pid == 0:                                         # rank r: all calls back to back, no synchronisation but the protocol's
            try:
                import signal
                signal.alarm(120)
                for n, which in enumerate(calls):
                    c.peer_allreduce_ll(parts[n][r], split, ptrs[which], nbytes, state[r], r, W, tmax, resid[r], resid[r], nw, 1e-6, xn[r], T)
                    log_h[r][n].copy_(resid[r])
                    log_n[r][n].copy_(xn[r])
                os._exit(0)
            except BaseException:
                import traceback
                traceback.print_exc()
                os._exit(1)
        pids.append(pid)
    codes = [os.waitpid(p, 0)[1] for p in pids]
    assert codes == [0] * W, codes
    cur = h0.clone()
    for n in range(len(calls)):
        cur = _expected(parts[n], split, cur, dt)
        x = cur.float()
        want = nw.float() * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).to(dt).float()
        for r in range(W):
            assert torch.equal(log_h[r][n], cur), f"call {n}: h of rank {r} differs from the rank-ordered sum"
            assert torch.equal(log_n[r][n], log_n[0][n]), f"call {n}: norm_out differs between ranks 0 and {r}"
        assert float((log_n[0][n].float() - want).abs().max() / want.abs().max()) < 1.2e-2
    assert [int(s[0]) for s in state] == [len(calls)] * W    # the epoch advanced once per call on every rank
