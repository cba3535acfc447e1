"""Iteration-level batching (chatts_b200/engine.py) on CPU through the C-ABI test double: requests join and leave the static
decode slots between graph replays; every request must produce exactly the tokens of a stand-alone greedy generate() call,
pages and slots must be recycled, and the engine must make progress with more requests than slots."""
import numpy as np
import torch

from tests.test_host_model import _build, _series


def _requests(proc, cfg):
    a, b = _series()
    specs = [("A <ts><ts/> and B <ts><ts/> ?", [a, b], 7), ("only text", [], 5), ("C <ts><ts/>", [b], 9), ("short", [], 1),
             ("another plain prompt that is a little longer", [], 6), ("D <ts><ts/> tail", [a[:90]], 4)]
    out = []
    for text, series, n in specs:
        enc = proc(text=[text], timeseries=series, return_tensors="pt")
        out.append((enc, n))
    return out


def test_requests_join_and_leave_and_match_standalone_generate(cabi_double):
    from chatts_b200.engine import ContinuousEngine
    cfg, sd, model, proc = _build(cabi_double)
    reqs = _requests(proc, cfg)
    ref = [model.generate(**enc, max_new_tokens=n, ignore_eos=True)[0, enc["input_ids"].shape[1]:].tolist() for enc, n in reqs]
    pages0 = len(model.pool.free)
    eng = ContinuousEngine(model, slots=2, steps_per_round=3, max_prefill_batch=2)
    for enc, n in reqs:
        eng.add_request(enc["input_ids"][0], enc["timeseries"], max_new_tokens=n, ignore_eos=True)
    done = eng.run()
    assert [r.rid for r in done] == list(range(len(reqs)))
    for r, want in zip(done, ref):
        assert r.tokens == want, (r.rid, r.tokens, want)
    assert max(eng.occupancy) == 2 and not eng.active and not eng.waiting       # never more than the slots, all drained
    assert len(model.pool.free) == pages0 - 1                                    # everything but the scratch page returned
    eng.close()
    assert len(model.pool.free) == pages0


def test_eos_frees_the_slot_early_and_late_arrivals_are_served(cabi_double):
    from chatts_b200.engine import ContinuousEngine
    cfg, sd, model, proc = _build(cabi_double)
    enc = proc(text=["plain prompt"], timeseries=[], return_tensors="pt")
    full = model.generate(**enc, max_new_tokens=8, ignore_eos=True)[0, enc["input_ids"].shape[1]:].tolist()
    eng = ContinuousEngine(model, slots=2, steps_per_round=2)
    rid = eng.add_request(enc["input_ids"][0], None, max_new_tokens=8, eos_token_id=[full[3]])     # its 4th token is "EOS"
    first = []
    while eng.has_work():
        first += eng.step()
    r = first[0]
    stop = full.index(full[3]) + 1
    assert r.rid == rid and r.tokens == full[:stop]
    # a request added after the engine went idle is served by the same engine / captured state
    eng.add_request(enc["input_ids"][0], None, max_new_tokens=5, ignore_eos=True)
    later = eng.run()
    assert later[0].tokens == full[:5]
    eng.close()
