"""Optional NVTX ranges around the phases of the hot path (the reference has no tracing at all).
Off by default -- ``CTS_NVTX=1`` turns them on so that an nsys / ncu --nvtx timeline shows prefill / TS encode / decode step /
training forward, loss, backward, optimiser as named ranges.  No effect on the launches themselves."""
import contextlib
import os

_ON = os.environ.get("CTS_NVTX", "0") == "1"


@contextlib.contextmanager
def span(name):
    if not _ON:
        yield
        return
    import torch
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def enabled():
    return _ON
