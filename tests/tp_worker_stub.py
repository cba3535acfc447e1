"""Stub engine for tests/test_vllm_tp_control.py: records the calls a spawned tensor-parallel rank receives (TEST INFRASTRUCTURE)."""
import json


class Recorder:
    def __init__(self, path):
        self.path = path

    def _generate(self, inputs, sp):
        with open(self.path, "a") as f:
            f.write(json.dumps({"prompts": [r["prompt"] for r in inputs], "max_tokens": sp.max_tokens, "temperature": sp.temperature}) + "\n")


def make_recorder(path):
    return Recorder(path)
