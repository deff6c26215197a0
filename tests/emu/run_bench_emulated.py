"""TEST INFRASTRUCTURE ONLY: runs bench.py's CUDA arm on a machine without a GPU, so that its control flow and the JSON contract
(value / e2e / roofline / cpu_baseline / clocks / gpu_launches ...) are exercised by the CPU suite (tests/test_emu.py).

    python tests/emu/run_bench_emulated.py --config 1 --steps 2 --warmup 3

Two substitutions are made in THIS process before bench.py's main() runs, neither of which exists in the product:
  * the package's C-ABI binding is pointed at the interpreter build of the kernels (tests/emu/_build/libvrt_cuda_emu.so);
  * the handful of torch.cuda facilities bench.py uses for plumbing (pinned / device tensors, streams, events, device
    properties) are replaced by host equivalents -- "device" tensors are host tensors, whose data_ptr() the interpreter's
    cudaMemcpy can read directly -- and, under torchrun, the NCCL process group by a gloo one (VRT_EMU_DEVICES >= ranks).
The numbers printed by such a run are meaningless as measurements (the interpreter is ~10^5 times slower than a B200) and are
never recorded; the test only checks the line's shape and the internal consistency of its fields.
"""
import contextlib
import os
import runpy
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE]


def install():
    import ctypes

    import torch

    import __graft_entry__ as ge
    import build_emu

    pkg = ge.load_package()
    lib = ctypes.CDLL(build_emu.build())
    for sym, (res, args) in pkg._ffi.CUDA_SYMBOLS.items():
        fn = getattr(lib, sym)
        fn.restype, fn.argtypes = res, args
    pkg._ffi._cuda = lib

    def host_only(fn):
        def wrapped(*a, **k):
            k.pop("pin_memory", None)
            if "device" in k and str(k["device"]).startswith("cuda"):
                k.pop("device")
            return fn(*a, **k)

        return wrapped

    for name in ("empty", "zeros", "tensor", "zeros_like", "full"):
        setattr(torch, name, host_only(getattr(torch, name)))

    class Event:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    class ExternalStream:
        def __init__(self, ptr, device=None):
            self.ptr = ptr

    # N > 1: gloo over host tensors in place of NCCL over device tensors (same collectives: broadcast, all_reduce, all_gather,
    # grouped isend/irecv, barrier)
    import torch.distributed as dist

    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **k: real_init("gloo")

    cuda = torch.cuda
    cuda.set_device = lambda d: None
    cuda.synchronize = lambda *a: None
    cuda.empty_cache = lambda: None
    cuda.Event = Event
    cuda.ExternalStream = ExternalStream
    cuda.stream = lambda s: contextlib.nullcontext()
    cuda.get_device_properties = lambda d: types.SimpleNamespace(multi_processor_count=int(os.environ.get("VRT_EMU_SMS", "2")), name="cuda_emu")


if __name__ == "__main__":
    install()
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
