"""Peer images (include/vrt_cuda.h: vrt_cuda_peer_image_create / _open / _close): a frame buffer on one GPU that another
PROCESS's render kernel stores its row band into -- the multi-GPU output path without a gather step.  The importing process may
sit on the same GPU (this test, so it runs on a one-GPU box) or on a peer (bench.py --gpus N)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CHILD = r"""
import sys
sys.path.insert(0, {root!r})
import __graft_entry__ as ge
pkg = ge.load_package()
V = pkg.vrt
r = V.Renderer(0)
ptr = r.peer_image_open(bytes.fromhex(sys.argv[1]))
W = 256
r.set_gaussians(pkg.scenes.grid(4))
cam, origin = V.camera_t.app(W, W)
flags = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
f = r.frame(cam.view_matrix, origin, W, W, flags, (16, 16), rows=(int(sys.argv[2]), int(sys.argv[3])))
r.tile(f)
r.render_device(f, ptr, 0)
r.sync()
r.peer_image_close(ptr)
r.close()
print("band stored")
"""


def test_another_process_renders_its_band_into_a_peer_image(pkg, renderer):
    if os.environ.get("VRT_EMU") == "1":
        pytest.skip("CUDA IPC needs two processes on real devices")
    import torch

    V = pkg.vrt
    W = 256
    renderer.set_gaussians(pkg.scenes.grid(4))
    cam, origin = V.camera_t.app(W, W)
    flags = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
    full = renderer.frame(cam.view_matrix, origin, W, W, flags, (16, 16))
    solo, _, _ = renderer.frame_render(full, True, False)

    ptr, handle = renderer.peer_image_create(W * W * 4)
    try:
        class Raw:
            __cuda_array_interface__ = {"shape": (W, W), "typestr": "<i4", "data": (ptr, False), "version": 2}

        image = torch.as_tensor(Raw(), device="cuda")
        image.zero_()
        torch.cuda.synchronize()
        # this process renders the upper band, the child the lower one, both into the same buffer
        top = renderer.frame(cam.view_matrix, origin, W, W, flags, (16, 16), rows=(0, 96))
        renderer.tile(top)
        renderer.render_device(top, ptr, 0)
        renderer.sync()
        out = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT), handle.hex(), "96", str(W)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "band stored" in out.stdout, out.stderr[-2000:]
        got = image.cpu().numpy().view(np.uint32)
        assert np.array_equal(got, solo)
        del image
    finally:
        renderer.peer_image_close(ptr)
    # a pointer the context does not own is refused, and so is a handle that names nothing
    with pytest.raises(V.VrtCudaError):
        renderer.peer_image_close(ptr)
    with pytest.raises(V.VrtCudaError):
        renderer.peer_image_open(bytes(64))
