import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# The three walking-v2 step kernels (DESIGN.md section 4 "Round 2").  The library picks one from N; the dynamics / invariant
# tests run small N, so they select each kernel explicitly and assert that it is the one that was launched.
WALK_KERNELS = {
    "two_warps_per_32_envs": ({}, "zbot_step_w2_kernel<"),
    "packed_halves": ({"ZBOT_W2": "0"}, "zbot_step_h2_kernel<"),
    "one_chain": ({"ZBOT_W2": "0", "ZBOT_H2": "0"}, "zbot_step_kernel<false"),
}


@pytest.fixture(params=list(WALK_KERNELS))
def walk_kernel(request, monkeypatch):
    for k in ("ZBOT_W2", "ZBOT_W2_CTAS", "ZBOT_H2", "ZBOT_STEP_VARIANT", "ZBOT_SWEEP_UNROLL"):
        monkeypatch.delenv(k, raising=False)
    env, prefix = WALK_KERNELS[request.param]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    return prefix
