"""CPU: the C-ABI library loads and exports every symbol include/chatts_b200.h declares; the product path
fails loudly (no fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "chatts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from chatts_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        from chatts_b200.build import build
        build(verbose=False)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/chatts_b200.h but not exported"
    assert sorted(_cabi.SYMBOLS) == syms, "chatts_b200/_cabi.py SYMBOLS out of sync with the header"
    lib.cts_arch.restype = ctypes.c_char_p
    assert lib.cts_arch() == b"sm_100a"
    assert lib.cts_version() >= 100


def test_sass_is_blackwell_native():
    """tcgen05 / TMA must be in the shipped SASS (UTCHMMA, UTMALDG, LDTM: B200_PROFILING.md 'what proves...')."""
    import shutil
    import subprocess
    from chatts_b200 import _cabi
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run([exe, "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_fallback_without_gpu():
    from chatts_b200 import ChatTSConfig, _cabi
    from chatts_b200.model import ChatTSForCausalLM
    from chatts_b200.weights import synthetic_state_dict
    with pytest.raises(_cabi.CtsError):
        _cabi.Context()
    cfg = ChatTSConfig.tiny()
    with pytest.raises(_cabi.CtsError):
        ChatTSForCausalLM(cfg, synthetic_state_dict(cfg, device="cpu"))


def test_product_never_imports_oracle():
    import glob
    for f in glob.glob(os.path.join(ROOT, "chatts_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
