import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge

    # build once per session if the in-tree libraries are missing (CPU container); the GPU box uses the prebuilt ones
    need = [os.path.join(ge.PKG_DIR, "csrc", n) for n in ("libvrt_cuda.so", "libvrt_host.so")] + [os.path.join(ROOT, "oracle", "libvrt_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def renderer(pkg):
    r = pkg.vrt.Renderer(0)
    yield r
    r.close()
