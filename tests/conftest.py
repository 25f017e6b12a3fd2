import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """A fresh checkout has no binaries: build the CUDA library (nvcc cross-compiles without a GPU) and the
    test-only plan interpreter once per session, like __graft_entry__.build() does."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "nhans_b200", "libnhans_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "nhans_b200", "csrc"), "-j8"])
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libplanexec.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    yield


def _has_gpu():
    try:
        from nhans_b200.engine import Engine
        e = Engine(0, 0, win_capacity=1, row_capacity=1)
        e.close()
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def weights_sn():
    from nhans_b200 import weights as W
    return W.seeded_init(W.SELECTIVE_NOISE, 0)


@pytest.fixture(scope="session")
def weights_ss():
    from nhans_b200 import weights as W
    return W.seeded_init(W.SEPARATOR, 0)


@pytest.fixture(scope="session")
def engine_sn(weights_sn):
    from nhans_b200.engine import Engine
    e = Engine(0, 0, win_capacity=256, row_capacity=4)     # must fail loudly without the CUDA library / a GPU
    e.load_weights(weights_sn)
    yield e
    e.close()


@pytest.fixture(scope="session")
def engine_ss(weights_ss):
    from nhans_b200.engine import Engine
    e = Engine(0, 1, win_capacity=256, row_capacity=4)
    e.load_weights(weights_ss)
    yield e
    e.close()


@pytest.fixture(scope="session")
def oracle_sn(weights_sn):
    from oracle import nhans_oracle as O
    return O.Net(weights_sn, 0)


@pytest.fixture(scope="session")
def oracle_ss(weights_ss):
    from oracle import nhans_oracle as O
    return O.Net(weights_ss, 1)
