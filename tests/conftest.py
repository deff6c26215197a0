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
    pkg_ = ge.load_package()
    if os.environ.get("VRT_EMU") == "1":
        _use_emulated_library(pkg_)
    return pkg_


def _use_emulated_library(pkg_):
    """VRT_EMU=1 (set only by tests/test_emu.py for its child pytest run): the `gpu` tests execute on the CPU against
    tests/emu/_build/libvrt_cuda_emu.so -- the product's CUDA sources compiled against the SIMT interpreter of
    tests/emu/cuda_emu.h.  Test infrastructure: the package itself never loads that library."""
    import ctypes

    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu

    # VRT_EMU_LIB: an alternative build of the same translation (e.g. -fsanitize=address, run under LD_PRELOAD=libasan.so)
    lib = ctypes.CDLL(os.environ.get("VRT_EMU_LIB") or build_emu.build())
    for sym, (res, args) in pkg_._ffi.CUDA_SYMBOLS.items():
        fn = getattr(lib, sym)
        fn.restype, fn.argtypes = res, args
    pkg_._ffi._cuda = lib
    # child processes (the volumetric-ray-tracer binary of tests/test_gpu_app.py) resolve the C ABI from the same library
    os.environ["LD_PRELOAD"] = ":".join(p for p in (os.environ.get("LD_PRELOAD"), lib._name) if p)


@pytest.fixture(scope="session", autouse=True)
def _emulation_for_tests_without_fixtures(request):
    """Under VRT_EMU=1 also the tests that only start the app binary need the library switch (LD_PRELOAD for children)."""
    if os.environ.get("VRT_EMU") == "1":
        request.getfixturevalue("pkg")


@pytest.fixture(scope="session")
def renderer(pkg):
    r = pkg.vrt.Renderer(0)
    yield r
    r.close()
