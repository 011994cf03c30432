import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def rel_err(a, b):
    """max|a-b| / max|b| — the parity metric of SURVEY.md §4 (tolerance 1e-3 from north_star)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_synthetic.pt"), weights_only=True)


@pytest.fixture(scope="session")
def weights():
    from lip2speech_b200 import spec
    return spec.seeded_state_dict(spec.full_spec(), 1234)


@pytest.fixture(scope="session")
def spk_weights(weights):
    p = "speaker_encoder."
    return {k[len(p):]: v for k, v in weights.items() if k.startswith(p)}
