"""OpenAI-style front end (chatts_b200/server.py) on CPU through the C-ABI test double: demo/vllm_api.py's request (text part +
{"timeseries": [...]} parts) verbatim through the `openai` client, the vLLM-shaped /v1/completions body, micro-batching of
concurrent requests, SSE streaming, and the reference's input errors as HTTP 400."""
import json
import threading

import numpy as np
import pytest

from tests.test_host_model import _build, _series


@pytest.fixture
def served(cabi_double):
    from starlette.testclient import TestClient
    from chatts_b200.server import create_app
    from chatts_b200.vllm_compat import LLM
    cfg, sd, model, proc = _build(cabi_double)
    app = create_app(LLM(model=model), batch_window_ms=150.0)
    with TestClient(app) as client:
        yield app, client
    app.state.engine.close()


def test_demo_vllm_api_request_through_the_openai_client(served):
    import openai
    app, http = served
    a, b = _series()
    prompt = "I have 2 time series. TS1: <ts><ts/>; TS2: <ts><ts/>. Compare them."
    prompt = f"<|im_start|>system\nYou are a helpful assistant.<|im_end|><|im_start|>user\n{prompt}<|im_end|><|im_start|>assistant\n"
    client = openai.OpenAI(base_url="http://testserver/v1", api_key="test", http_client=http)
    r = client.chat.completions.create(model="chatts", max_tokens=7, extra_body={"ignore_eos": True},
                                       messages=[{"role": "user", "content": [{"type": "text", "text": prompt}] +
                                                  [{"timeseries": ts} for ts in (a.tolist(), b.tolist())]}])
    assert r.choices[0].message.role == "assistant" and isinstance(r.choices[0].message.content, str)
    assert r.usage.completion_tokens == 7 and r.choices[0].finish_reason == "length"
    assert [m.id for m in client.models.list().data] == ["chatts"]


def test_completions_body_batching_and_errors(served):
    app, http = served
    a, _ = _series()
    body = {"prompt": "x <ts><ts/> y", "multi_modal_data": {"timeseries": [a.tolist()]}, "max_tokens": 5, "ignore_eos": True}
    ref = http.post("/v1/completions", json=body).json()
    assert ref["choices"][0]["finish_reason"] == "length" and ref["usage"]["completion_tokens"] == 5
    # concurrent identical-parameter requests are decoded as ONE batch, and each caller gets its own answer back
    res = [None] * 4
    def call(i):
        bd = dict(body) if i % 2 == 0 else {"prompt": f"plain text {i}", "max_tokens": 5, "ignore_eos": True}
        res[i] = http.post("/v1/completions", json=bd).json()
    n0 = len(app.state.engine.batches)
    ths = [threading.Thread(target=call, args=(i,)) for i in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert all(r["usage"]["completion_tokens"] == 5 for r in res)
    assert sum(app.state.engine.batches[n0:]) == 4              # (how they were grouped depends on arrival times: see the engine test)
    assert res[0]["choices"][0]["text"] == ref["choices"][0]["text"] == res[2]["choices"][0]["text"]    # greedy: batch-invariant here
    # the reference's input errors -> 400
    assert http.post("/v1/completions", json={"prompt": "two <ts><ts/> <ts><ts/>", "multi_modal_data": {"timeseries": [a.tolist()]}}).status_code == 400
    assert http.post("/v1/chat/completions", json={"messages": [{"role": "user", "content": [{"image_url": "x"}]}]}).status_code == 400
    many = {"prompt": "<ts><ts/>" * 16, "multi_modal_data": {"timeseries": [a.tolist()] * 16}}
    assert http.post("/v1/completions", json=many).status_code == 400                               # --limit-mm-per-prompt timeseries=15
    assert http.get("/health").json() == {"status": "ok"}


def test_streaming_chunks_concatenate_to_the_full_answer(served):
    app, http = served
    msgs = [{"role": "system", "content": "You are a helpful assistant."}, {"role": "user", "content": "hello there"}]
    full = http.post("/v1/chat/completions", json={"messages": msgs, "max_tokens": 6, "ignore_eos": True}).json()
    pieces, done = [], False
    with http.stream("POST", "/v1/chat/completions", json={"messages": msgs, "max_tokens": 6, "ignore_eos": True, "stream": True}) as r:
        for line in r.iter_lines():
            if not line.startswith("data: "):
                continue
            if line == "data: [DONE]":
                done = True
                break
            d = json.loads(line[6:])["choices"][0]["delta"]
            pieces.append(d.get("content", ""))
    assert done and "".join(pieces) == full["choices"][0]["message"]["content"] and len([p for p in pieces if p != ""]) <= 6


def test_messages_to_prompt_chatml_and_parts():
    from chatts_b200.server import messages_to_prompt
    p, s = messages_to_prompt([{"role": "system", "content": "S"}, {"role": "user", "content": [{"type": "text", "text": "a <ts><ts/>"},
                                                                                                 {"timeseries": [1, 2, 3]}]}])
    assert p == "<|im_start|>system\nS<|im_end|><|im_start|>user\na <ts><ts/><|im_end|><|im_start|>assistant\n" and s == [[1, 2, 3]]
    raw = "<|im_start|>user\nhi<|im_end|><|im_start|>assistant\n"
    assert messages_to_prompt([{"role": "user", "content": raw}])[0] == raw


def test_continuous_scheduler_serves_concurrent_requests(cabi_double):
    """--scheduler continuous: more concurrent requests than slots, different lengths; each answer equals the stand-alone
    greedy generation; a streamed request receives the same text piecewise; sampling is refused with 400."""
    from starlette.testclient import TestClient
    from chatts_b200.server import create_app
    from chatts_b200.vllm_compat import LLM, SamplingParams
    cfg, sd, model, proc = _build(cabi_double)
    llm = LLM(model=model)
    a, _ = _series()
    bodies = [{"prompt": "x <ts><ts/> y", "multi_modal_data": {"timeseries": [a.tolist()]}, "max_tokens": 6, "ignore_eos": True}] + \
             [{"prompt": f"plain text number {i}", "max_tokens": 3 + i, "ignore_eos": True} for i in range(9)]
    want = []
    for b in bodies:
        req = {"prompt": b["prompt"], "multi_modal_data": b.get("multi_modal_data", {})}
        want.append(llm.generate([req], SamplingParams(max_tokens=b["max_tokens"], ignore_eos=True))[0].outputs[0].text)
    app = create_app(llm, scheduler="continuous", steps_per_round=2)
    with TestClient(app) as http:
        res = [None] * len(bodies)
        def call(i):
            res[i] = http.post("/v1/completions", json=bodies[i]).json()
        ths = [threading.Thread(target=call, args=(i,)) for i in range(len(bodies))]
        [t.start() for t in ths]
        [t.join() for t in ths]
        assert [r["choices"][0]["text"] for r in res] == want
        assert max(app.state.engine.occupancy) <= model.max_batch
        pieces = []
        with http.stream("POST", "/v1/completions", json={**bodies[3], "stream": True}) as r:
            for line in r.iter_lines():
                if line.startswith("data: ") and line != "data: [DONE]":
                    pieces.append(json.loads(line[6:])["choices"][0]["text"])
        assert "".join(pieces) == want[3]
        assert http.post("/v1/completions", json={"prompt": "x", "temperature": 0.7}).status_code == 400
    app.state.engine.close()
    assert len(model.pool.free) == model.pool.num_pages


@pytest.mark.parametrize("scheduler", ["batch", "continuous"])
def test_engine_groups_queued_requests_deterministically(cabi_double, scheduler):
    """Requests queued before the worker starts: the batch scheduler decodes equal-parameter requests as ONE batch (and a
    request with other parameters on its own), the continuous scheduler fills its slots -- independent of thread timing."""
    from chatts_b200.server import Engine
    from chatts_b200.vllm_compat import LLM
    cfg, sd, model, proc = _build(cabi_double)
    eng = Engine(LLM(model=model), batch_window_ms=50.0, scheduler=scheduler, steps_per_round=2, autostart=False)
    p5 = {"max_tokens": 5, "temperature": 0.0, "top_p": 1.0, "top_k": 0, "n": 1, "ignore_eos": True}
    jobs = [eng.submit(f"prompt {i}", [], p5) for i in range(3)] + [eng.submit("other", [], {**p5, "max_tokens": 3})]
    eng.start()
    outs = [j.future.result(timeout=120) for j in jobs]
    assert [len(o.outputs[0].token_ids) for o in outs] == [5, 5, 5, 3]
    if scheduler == "batch":
        assert eng.batches == [3, 1]
    else:
        assert max(eng.occupancy) == 4
    eng.close()
