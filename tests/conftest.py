import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def tiny_engine():
    """A small context for operator-level tests (no weights needed for the ccb_op_* / sampler entry points)."""
    import torch
    import clipcap_b200 as cc
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = cc.EngineConfig(lm_d=128, lm_layers=1, lm_heads=2, lm_vocab=503, lm_n_pos=64, map_kind="none", vit=False,
                          max_images=8, max_beam=5, max_ctx=32)
    eng = cc.Engine(cfg)
    yield eng
    eng.close()
