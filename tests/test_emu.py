"""CPU execution of the product's CUDA sources under the SIMT interpreter of tests/emu (TEST INFRASTRUCTURE).

The container that runs `pytest -m "not gpu"` has no GPU, so without this file the kernels and their launch logic would
only ever be exercised on the B200 box.  tests/emu/build_emu.py compiles a rewritten COPY of
simd-gaussian-ray-tracing_b200/csrc/{vrt_cuda.cu,*.cuh} against tests/emu/cuda_emu.h (every CUDA thread a fiber, warp
collectives and __syncthreads as rendezvous, guarded device allocations, an mbarrier / bulk-copy model) into
tests/emu/_build/libvrt_cuda_emu.so, and this file runs a selection of the `gpu` parity tests against that library in a
child pytest (VRT_EMU=1, see tests/conftest.py) -- the same test functions, the same oracle, the same tolerances, through
the same C ABI.  The interpreter aborts on divergent collectives, deadlocks, __trap, writes outside an allocation and
bulk copies nobody waits for, so those show up here as failures too.

This is not a product path and not a fallback: nothing under simd-gaussian-ray-tracing_b200/ can load the emulated
library (the package binds csrc/libvrt_cuda.so only and fails without a GPU), and numbers produced here are never
reported as measurements.  Arithmetic differs from the GPU by a few ulp (no MUFU approximations, no FMA contraction), so
the GPU run at round end stays the parity gate; this run pins the LOGIC (indexing, lists, queues, bands, split cells,
staging) at small sizes.
"""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

# gpu tests small enough for the interpreter (a frame of 256 x 256 pixels takes about a second); the full-size frames
# of BASELINE configs 2-5 and the throughput probes stay GPU-only
EMU_SELECTION = [
    "tests/test_gpu_parity.py::test_config1_tiled",
    "tests/test_gpu_parity.py::test_config1_untiled",
    "tests/test_gpu_parity.py::test_config1_against_compiled_reference",
    "tests/test_gpu_parity.py::test_membership_config1",
    "tests/test_gpu_parity.py::test_membership_img_error_scene",
    "tests/test_gpu_parity.py::test_membership_grid64_and_rotated",
    "tests/test_gpu_parity.py::test_membership_objects",
    "tests/test_gpu_parity.py::test_row_bands_compose",
    "tests/test_gpu_parity.py::test_bound_mode_matches_all",
    "tests/test_gpu_parity.py::test_bounded_lists_are_conservative_and_tight",
    "tests/test_gpu_parity.py::test_errors_are_reported",
    "tests/test_gpu_parity.py::test_empty_scene_and_ragged_image",
    "tests/test_gpu_parity.py::test_scene_out_of_view_renders_black_in_every_list_mode",
    "tests/test_gpu_approx.py::test_tables_match_reference",
    "tests/test_gpu_approx.py::test_variants_on_device_built_lists",
    "tests/test_gpu_approx.py::test_variant_errors",
    "tests/test_gpu_approx.py::test_tables_against_compiled_reference_live",
    # the heavier paths at interpreter-friendly sizes: depth window, split cells, bulk-copy staging, kernel variants
    "tests/test_gpu_small_frames.py",
    # the callers either side of the path: the volumetric-ray-tracer binary and the C++ drop-in check built against the
    # reference's own types resolve the C ABI from the interpreter build through LD_PRELOAD (tests/conftest.py)
    "tests/test_gpu_app.py",
]
# gpu tests that also pass under the interpreter but take minutes there (VRT_EMU_FULL=1 adds them): OBJ scenes, the
# img-error procedure, thin bands of a dense frame, the 512^2 depth-window frames, the NO_SKIP walk of the monkey
EMU_SLOW = [
    "tests/test_gpu_parity.py::test_objects_tiled",
    "tests/test_gpu_parity.py::test_teapot_subsample",
    "tests/test_gpu_parity.py::test_host_tile_lists_equal_device_lists",
    "tests/test_gpu_parity.py::test_thin_bands_of_a_dense_frame_compose_bit_exactly",
    "tests/test_gpu_parity.py::test_depth_window_mode_is_the_same_image",
]


@pytest.fixture(scope="module")
def build_emu():
    import build_emu as b

    return b


def test_interpreter_selftest(build_emu):
    """Known answers for the collectives / barrier / atomics / bulk-copy model, and the interpreter's own checks: each of
    the deliberately broken kernels of tests/emu/selftest.cu must abort with its diagnosis."""
    lib = build_emu.build_selftest()
    assert ctypes.CDLL(lib).emu_selftest(0) == 0
    expect = {1: "write outside a device allocation", 2: "divergent collectives", 3: "deadlock", 4: "never waited for"}
    for case, message in expect.items():
        r = subprocess.run([sys.executable, "-c", f"import ctypes; ctypes.CDLL({lib!r}).emu_selftest({case})"], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and message in r.stderr, (case, r.returncode, r.stderr[-500:])
    # a missing __syncwarp is invisible in one lane order and shows in the other (after a collective the last lane to arrive runs
    # on first: lane 31 in the forward order, lane 0 in the reversed one), so every selection is run in both orders
    seen = {}
    for order in ("forward", "reverse"):
        r = subprocess.run([sys.executable, "-c", f"import ctypes, sys; sys.exit(ctypes.CDLL({lib!r}).emu_selftest(5))"], env=dict(os.environ, VRT_EMU_ORDER=order),
                           capture_output=True, text=True, timeout=120)
        seen[order] = r.returncode
    assert max(seen.values()) == 32 and min(seen.values()) < 32, seen


def test_translation_covers_every_cuda_construct(build_emu):
    """The rewrite must account for every launch and every inline-PTX statement of the product sources (it raises on
    anything it does not know), and must leave the product tree untouched."""
    csrc = build_emu.CSRC
    n_launch = n_asm = 0
    for f in build_emu.sources():
        src = open(os.path.join(csrc, f)).read()
        out = build_emu.translate(f, src)
        n_launch += src.count("<<<")
        n_asm += len(__import__("re").findall(r"\basm\b", src))
        assert out.count("emu::launch(") == src.count("<<<"), f
    assert n_launch >= 30 and n_asm >= 8
    assert not any("emu" in f.lower() for f in os.listdir(csrc)), "emulation artefacts do not belong in the product tree"


def test_package_cannot_load_the_emulated_library(pkg):
    """The product binding names csrc/libvrt_cuda.so and nothing else."""
    src = open(os.path.join(ROOT, "simd-gaussian-ray-tracing_b200", "_ffi.py")).read()
    assert "libvrt_cuda.so" in src and "emu" not in src.lower()
    assert os.path.realpath(pkg._ffi.cuda_lib()._name) == os.path.realpath(os.path.join(ROOT, "simd-gaussian-ray-tracing_b200", "csrc", "libvrt_cuda.so"))


@pytest.mark.parametrize("tma", ["lazy", "eager"])
def test_gpu_parity_tests_under_the_interpreter(build_emu, tma):
    """The selected `gpu` tests, unchanged, against the emulated library.  Bulk copies complete either at the first wait on
    their mbarrier (lazy: a consumer that forgets to wait reads stale data) or at issue (eager: a producer that refills a
    buffer still being read clobbers it); the TMA-staged paths must pass under both."""
    build_emu.build()
    selection = EMU_SELECTION if tma == "lazy" else ["tests/test_gpu_small_frames.py::test_contiguous_lists_through_the_bulk_copy_staging", "tests/test_gpu_parity.py::test_config1_untiled"]
    if tma == "lazy" and os.environ.get("VRT_EMU_FULL") == "1":
        selection = selection + EMU_SLOW
    env = dict(os.environ, VRT_EMU="1", VRT_EMU_TMA=tma, VRT_EMU_DEVICES="3")  # three pretend GPUs for the app's --gpus test
    if tma == "eager":
        # the second leg also runs the lanes of every warp from 31 down to 0: shared memory handed from one lane to another
        # without a __syncwarp (or a copy waited for before lane 0 issued it) only works in the forward order
        env["VRT_EMU_ORDER"] = "reverse"
        selection = selection + ["tests/test_gpu_small_frames.py::test_depth_window_small", "tests/test_gpu_small_frames.py::test_split_cells_and_bands_small",
                                 "tests/test_gpu_parity.py::test_row_bands_compose", "tests/test_gpu_approx.py::test_variants_on_device_built_lists"]
    # (the randomised sweep runs in the AddressSanitizer test below, with fewer cases)
    # (the long-list test runs with every long list kept in k2_band_long; its second leg, K1's default wide mark, is left to
    # VRT_EMU_FULL and the GPU: 25 s under the interpreter)
    extra = [] if os.environ.get("VRT_EMU_FULL") == "1" else ["--deselect", "tests/test_gpu_small_frames.py::test_long_lists_share_one_cache_per_cta[0.4]"]
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "--deselect",
                        "tests/test_gpu_small_frames.py::test_randomised_small_frames", *extra, *selection], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout, tail


def test_kernels_under_address_sanitizer(build_emu):
    """The same translation built with -fsanitize=address: device buffers are plain heap blocks there, so a kernel (or the
    launch logic) that reads or writes one element past a list, a queue, a record array or the image aborts with the source
    line.  Ragged images, empty scenes, row bands, split cells, the depth window and the bulk-copy staging are the index
    arithmetic most likely to be off by one."""
    asan = build_emu.libasan()
    if asan is None:
        pytest.skip("this g++ has no libasan.so")
    lib = build_emu.build(asan=True)
    small = "tests/test_gpu_small_frames.py::"
    selection = [small + "test_split_cells_and_bands_small", small + "test_contiguous_lists_through_the_bulk_copy_staging",
                 small + "test_k1_bins_cells_and_bands", "tests/test_gpu_parity.py::test_empty_scene_and_ragged_image", "tests/test_gpu_parity.py::test_row_bands_compose",
                 "tests/test_gpu_parity.py::test_scene_out_of_view_renders_black_in_every_list_mode"]
    if os.environ.get("VRT_EMU_FULL") == "1":
        selection += [small + "test_depth_window_small", small + "test_depth_window_long_lists_take_the_in_loop_test", small + "test_long_lists_share_one_cache_per_cta", small + "test_register_block_and_packing_variants_small",
                      "tests/test_gpu_parity.py::test_config1_untiled", "tests/test_gpu_approx.py::test_variants_on_device_built_lists"]
    env = dict(os.environ, VRT_EMU="1", VRT_EMU_LIB=lib, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", *selection], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-3000:]
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    # randomised frames (tests/emu/fuzz_frames.py): odd image sizes, list lengths around the 32-record staging chunks, every
    # list mode, NO_SKIP / depth window / pinned slice / emitter block, arbitrary row bands -- against the oracle, under ASan
    cases = "400" if os.environ.get("VRT_EMU_FULL") == "1" else "40"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "fuzz_frames.py"), "--emu", "--cases", cases, "--seed", "3"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "fuzz ok" in r.stdout and "AddressSanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-3000:]
    # lists beyond k2_band's per-warp cache, every one of them kept in k2_band_long (VRT_CUDA_LONG_WIDE=2: K1 marks none as wide),
    # then with the default mark (most of these wide-sigma lists then take k2_render<WIN>)
    for wide in ("2", "0.4"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "fuzz_frames.py"), "--emu", "--cases", "0", "--long-cases", "200" if cases == "400" else "10",
                            "--seed", "7"], cwd=ROOT, env=dict(env, VRT_CUDA_LONG_WIDE=wide), capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0 and "fuzz ok" in r.stdout and "AddressSanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-3000:]


def _rank_worker(rank, world, port, lib_path, result_path):
    """One rank of bench.py's multi-GPU protocol with the emulated library standing in for the rank's GPU and gloo for NCCL."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    pkg = ge.load_package()
    lib = ctypes.CDLL(lib_path)
    for sym, (res, args) in pkg._ffi.CUDA_SYMBOLS.items():
        fn = getattr(lib, sym)
        fn.restype, fn.argtypes = res, args
    pkg._ffi._cuda = lib
    from vrt_b200 import bands as B

    V = pkg.vrt
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, tiles = 96, 12
    # rank 0 owns the scene; everyone receives it (bench.py: dist.broadcast of the device copy)
    scene = torch.from_numpy(pkg.scenes.synthetic(1500, 5, -1.6, -1.2)) if rank == 0 else torch.zeros((1500, 10), dtype=torch.float32)
    dist.broadcast(scene, 0)
    r = V.Renderer(0)
    r.set_gaussians(scene.numpy())
    cam, origin = V.camera_t.app(W, W, rotation=9.0)
    flags = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
    full = r.frame(cam.view_matrix, origin, W, W, flags, (tiles, tiles), 6.0)
    r.tile(full)
    row_cost, row_px = r.row_costs()
    bounds = B.split_rows(row_cost, row_px, world, W, align=W // tiles)
    r.set_slice(r.auto_slice(1.0 / world))
    band = r.frame(cam.view_matrix, origin, W, W, flags, (tiles, tiles), 6.0, rows=(bounds[rank], bounds[rank + 1]))
    image = np.zeros((W, W), np.uint32)
    r.tile(band)
    _, _, st = r.render(band, True, False, image=image)
    t = torch.from_numpy(image.view(np.int32))
    B.gather_bands(t, bounds, rank, world, dist)
    terms = torch.tensor([st["terms_listed"]], dtype=torch.float64)
    dist.all_reduce(terms)
    if rank == 0:
        whole, _, st_full = r.frame_render(full, True, False)
        with open(result_path, "w") as f:
            f.write(f"{int(np.array_equal(whole, image))} {int(terms.item() == st_full['terms_listed'])} {int(whole.any())} {' '.join(map(str, bounds))}")
    dist.barrier()
    dist.destroy_process_group()
    r.close()


@pytest.mark.parametrize("world", [2, 3])
def test_multi_rank_protocol_with_rendering(build_emu, world, tmp_path):
    """bench.py's N > 1 protocol end to end on the CPU: scene broadcast, full-frame tiling for the row costs, cost-balanced
    bands on tile rows, one slice size for every rank, every rank renders its band with the real kernels (interpreter), bands
    gathered to rank 0 (gloo in place of NCCL) -- and the gathered image must equal the frame one context renders alone, bit
    for bit, with the listed work of the bands adding up to the frame's."""
    import torch.multiprocessing as mp
    from test_bands import _free_port

    result = str(tmp_path / "res.txt")
    mp.spawn(_rank_worker, args=(world, _free_port(), build_emu.build(), result), nprocs=world, join=True)
    equal, terms_add_up, drawn, *bounds = open(result).read().split()
    assert (equal, terms_add_up, drawn) == ("1", "1", "1"), (equal, terms_add_up, drawn, bounds)
    assert len(bounds) == world + 1 and bounds[0] == "0" and bounds[-1] == "96"


@pytest.mark.parametrize("extra", [[], ["--lists", "reference", "--slice", "16", "--no-cpu-baseline", "--only"]])
def test_bench_cuda_arm_contract(build_emu, extra):
    """bench.py's own arm cannot run without a GPU, so its control flow and the JSON line the driver parses are exercised here
    with the interpreter standing in for the device and host tensors for torch's device tensors
    (tests/emu/run_bench_emulated.py).  The numbers are meaningless as measurements; the line's shape and the consistency of
    its fields are what is checked."""
    import json

    build_emu.build()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_bench_emulated.py"), "--config", "1", "--configs", "0", "--steps", "2", "--warmup", "3", *extra], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "gpu_launches", "clocks", "roofline"):
        assert key in d, key
    assert "impl" not in d and d["metric"] == "pixel-Gaussian evaluations/s" and d["unit"] == "evals/s" and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    cfg = d["config"]
    assert cfg["workload"].startswith("config1") and "model" not in cfg and cfg["bands"] == [0, 256]
    assert 0 < cfg["terms_evaluated_per_frame"] <= cfg["terms_resolved_per_frame"] <= cfg["terms_listed_per_frame"]
    assert abs(cfg["terms_resolved_per_frame"] - (cfg["terms_evaluated_per_frame"] + cfg["terms_saturated_per_frame"] + cfg["terms_terminated_per_frame"])) <= 1e-9 * cfg["terms_resolved_per_frame"]
    # value = resolved terms / time of the timed steps; the strict figure counts the evaluated ones only
    assert abs(d["value"] - cfg["terms_resolved_per_frame"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert abs(d["strict_evals_per_s"] - cfg["terms_evaluated_per_frame"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    e2e = d["e2e"]
    assert e2e["unit"] == d["unit"] and e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 16 * 40 and e2e["d2h_bytes_per_step"] == 256 * 256 * 4
    assert "vrt_cuda_render(host image)" in e2e["path"] and d["e2e_pageable"]["value"] > 0  # N = 1: the reference-facing host-buffer call
    assert d["gpu_launches"] > 0
    roof = d["roofline"]
    assert roof["bound"] == "fp32" and roof["kernel"] == "k2_band" and roof["unit"] == "TFLOP/s" and roof["flops_per_term"] == 15.0
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) <= 1e-9
    assert abs(roof["achieved"] - roof["terms_per_launch"] * 15.0 / (roof["ms_per_launch"] * 1e-3) / 1e12) <= 1e-6 * roof["achieved"]
    strict = cfg["strict_mode"]
    sroof = strict["roofline"]
    assert sroof["kernel"] == "k2_render" and strict["ms_per_step"] > 0 and strict["terms_evaluated_per_frame"] >= cfg["terms_evaluated_per_frame"]
    pm = sroof["pipe_model"]
    assert pm["ceiling_terms_per_s"] == 2 * 4 * 32 / 8.0 * 1965e6 and 0 < pm["frac_of_ceiling"] < 1  # 2 pretend SMs at the B200's max clock
    assert abs(pm["frac_of_ceiling"] / sroof["frac"] - sroof["peak"] * 1e12 / 15.0 / pm["ceiling_terms_per_s"]) <= 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if extra:
        assert "cpu_baseline" not in d and cfg["slice"] == 16 and "literal" in cfg["lists"] and "configs" not in d
        assert cfg["terms_listed_per_frame"] == 256 * 256 * 5 * 81  # the reference's own lists: 9 Gaussians in every tile
    else:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == "evals/s" and cb["sample"]
        assert d["parity_vs_reference"] is None or d["parity_vs_reference"]["max_channel_lsb"] <= 1  # config 1 is well conditioned
        assert set(d["configs"]) == {"0", "1"} and d["configs"]["0"]["ms_per_frame"] > 0 and d["configs"]["0"]["roofline_frac"] > 0


def test_bench_cuda_arm_two_ranks(build_emu):
    """The driver's N > 1 launch line (torch.distributed.run, one rank per GPU) with two interpreter instances and gloo in
    place of NCCL: only rank 0 prints, the bands cover the image, and the gathered image equals the frame one rank renders."""
    import json

    from test_bands import _free_port

    build_emu.build()
    env = dict(os.environ, VRT_EMU_DEVICES="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "emu", "run_bench_emulated.py"), "--gpus", "2", "--config", "1", "--configs", "0", "--steps", "2", "--warmup", "3"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    cfg = d["config"]
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and "cpu_baseline" not in d
    assert len(cfg["bands"]) == 3 and cfg["bands"][0] == 0 and cfg["bands"][-1] == 256 and all(b % 16 == 0 for b in cfg["bands"])
    assert cfg["gathered_image_equals_single_gpu_frame"] is True
    assert len(cfg["per_rank"]["ms_render"]) == 2 and abs(sum(cfg["per_rank"]["terms_executed"]) - cfg["terms_evaluated_per_frame"]) <= 1e-9 * cfg["terms_evaluated_per_frame"]
    assert d["e2e"]["value"] > 0 and d["roofline"]["frac"] > 0
