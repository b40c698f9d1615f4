import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "7bgzf_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # test infrastructure (generator, emulator, oracle) is plain C/C++: build it on demand
    need = ["build/libdatagen.so", "build/libemul.so", "oracle/liboracle.so"]
    if not all(os.path.exists(os.path.join(ROOT, p)) for p in need):
        subprocess.run(["make", "-s", "testlibs"], cwd=ROOT, check=True)


@pytest.fixture(scope="session")
def root():
    return ROOT
