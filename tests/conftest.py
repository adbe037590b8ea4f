import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the oracle is test infrastructure: (re)build it when a compiler is around
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                          env={k: v for k, v in os.environ.items() if k != "LD_PRELOAD"})      # (a sanitizer runtime preloaded for the test process is not for make)


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    return _oracle
