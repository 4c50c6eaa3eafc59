"""pytest configuration: `gpu` marker (tests that need a B200), shared synthetic fixtures."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def v2_ctc_ckpt():
    from gigaam_b200 import synthetic
    return synthetic.synthetic_checkpoint("v2_ctc", seed=0)


@pytest.fixture(scope="session")
def v2_rnnt_ckpt():
    from gigaam_b200 import synthetic
    return synthetic.synthetic_checkpoint("v2_rnnt", seed=0)


@pytest.fixture(scope="session")
def v1_ctc_ckpt():
    from gigaam_b200 import synthetic
    return synthetic.synthetic_checkpoint("v1_ctc", seed=0)


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
