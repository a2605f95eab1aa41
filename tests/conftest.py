import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oraclelib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oraclelib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libgpc_ref.so not present (built only where /root/reference exists)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def small_cases():
    z = np.load(os.path.join(ROOT, "tests", "golden", "small_cases.npz"))
    names = sorted({k.split("_")[0] for k in z.files})
    return {n: {k.split("_", 1)[1]: z[k] for k in z.files if k.startswith(n + "_")} for n in names}
