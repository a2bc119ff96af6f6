import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-size configuration, minutes of CPU time")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """The CPU tier needs the oracle, the scene compiler and the test-only emulation library."""
    import subprocess

    need = ["polaris_b200/libpolaris_scene.so", "oracle/libpolaris_oracle.so", "tests/emul/libpc_emul.so", "tests/emul/libpc_emul_popcull.so",
            "polaris_b200/libpolaris_cuda.so"]
    if not all(os.path.exists(os.path.join(ROOT, p)) for p in need):
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)
