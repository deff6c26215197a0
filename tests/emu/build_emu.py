"""TEST INFRASTRUCTURE ONLY -- builds tests/emu/_build/libvrt_cuda_emu.so: the product's CUDA sources compiled for the CPU
against the SIMT interpreter of cuda_emu.h, so the CPU test suite can execute the real kernels and launch logic.

The product sources are not modified and carry no emulation hooks: this script rewrites a COPY of
simd-gaussian-ray-tracing_b200/csrc/{vrt_cuda.cu,*.cuh} into tests/emu/_build/:
    #include <cuda_runtime.h>                         -> #include "cuda_emu.h"
    kernel<<<grid, block, smem, stream>>>(args)        -> emu::launch([&]() { kernel(args); }, grid, block, smem, stream)
    asm("ex2.approx.ftz.f32 ..."), rcp, mbarrier.*,
    cp.async.bulk, fence.*                             -> emu:: helpers with the same operands
    extern __shared__ ... s_raw[];                     -> unsigned char *s_raw = emu::dynamic_smem();
Every rewrite asserts that it matched, so a change of the product sources that the script does not understand fails the
build instead of silently testing something else.  Nothing in the package loads the result; only tests/test_emu.py does.
"""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "simd-gaussian-ray-tracing_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libvrt_cuda_emu.so")


def _match_close(s, i, open_c, close_c):
    """index of the bracket closing the one at s[i]"""
    depth = 0
    while True:
        c = s[i]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1


def rewrite_launches(src):
    out, i, n = [], 0, 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out.append(src[i:])
            return "".join(out), n
        k = j
        while src[k - 1].isspace():
            k -= 1
        if src[k - 1] == ">":  # template arguments of the kernel
            depth = 0
            while True:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while src[k - 1].isalnum() or src[k - 1] in "_:":
            k -= 1
        kernel = src[k:j].strip()
        e = src.index(">>>", j)
        cfg = src[j + 3 : e]
        p = src.index("(", e + 3)
        assert src[e + 3 : p].strip() == "", f"unexpected text between >>> and the argument list: {src[e:p+1]!r}"
        q = _match_close(src, p, "(", ")")
        args = src[p + 1 : q]
        out.append(src[i:k])
        out.append(f"emu::launch([&]() {{ {kernel}({args}); }}, {cfg})")
        i = q + 1
        n += 1


_OPERAND = re.compile(r'"[=+]?[rlfh]"\s*\(')


def _asm_operands(body):
    ops, i = [], 0
    while True:
        m = _OPERAND.search(body, i)
        if not m:
            return ops
        p = m.end() - 1
        q = _match_close(body, p, "(", ")")
        ops.append(body[p + 1 : q].strip())
        i = q + 1


def _asm_text(body):
    """the instruction template: the string literals before the first ':' that is outside a literal"""
    parts, i = [], 0
    while i < len(body):
        c = body[i]
        if c == '"':
            j = i + 1
            while body[j] != '"':
                j += 2 if body[j] == "\\" else 1
            parts.append(body[i + 1 : j])
            i = j + 1
        elif c == ":":
            break
        else:
            i += 1
    return "".join(parts)


def rewrite_asm(src):
    out, i, n = [], 0, 0
    pat = re.compile(r"\basm\s*(?:volatile)?\s*\(")
    while True:
        m = pat.search(src, i)
        if not m:
            out.append(src[i:])
            return "".join(out), n
        p = m.end() - 1
        q = _match_close(src, p, "(", ")")
        semi = src.index(";", q)
        body = src[p + 1 : q]
        text = _asm_text(body)
        ops = _asm_operands(body)
        if "ex2.approx.ftz.f32" in text:
            rep = f"{ops[0]} = emu::ex2_ftz({ops[1]});"
        elif "rcp.approx.ftz.f32" in text:
            rep = f"{ops[0]} = emu::rcp_ftz({ops[1]});"
        elif "mbarrier.init" in text:
            rep = f"emu::mbar_init({ops[0]}, {ops[1]});"
        elif "mbarrier.try_wait.parity" in text:
            rep = f"{ops[0]} = emu::mbar_try_wait({ops[1]}, {ops[2]});"
        elif "mbarrier.arrive.expect_tx" in text:
            rep = f"emu::mbar_expect_tx({ops[0]}, {ops[1]});"
        elif "cp.async.bulk" in text:
            rep = f"emu::bulk_copy({ops[0]}, {ops[1]}, {ops[2]}, {ops[3]});"
        elif "fence.proxy.async" in text or "fence.mbarrier_init" in text:
            rep = ";"
        else:
            raise AssertionError(f"build_emu.py does not know this inline PTX: {text!r}")
        out.append(src[i : m.start()])
        out.append(rep)
        i = semi + 1
        n += 1


def translate(name, src):
    src, n_launch = rewrite_launches(src)
    src, n_asm = rewrite_asm(src)
    # dynamic shared memory: one buffer per CTA, handed out by the interpreter
    src, n_dyn = re.subn(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];", r"unsigned char *\1 = emu::dynamic_smem();", src)
    if name == "k2_long.cuh":
        assert n_dyn == 1, "k2_long.cuh: the dynamic shared memory declaration was not found"
    if name.endswith(".cu"):
        assert src.count("#include <cuda_runtime.h>") == 1
        src = src.replace("#include <cuda_runtime.h>", '#include "cuda_emu.h"')
    if name == "vrt_cuda.cu":
        assert n_launch >= 30, f"only {n_launch} kernel launches found in vrt_cuda.cu"
        # The interpreter runs one kernel at a time on process-wide state and keeps ONE copy of the __constant__ frame geometry
        # (a GPU has one per device): every C-ABI entry that touches device state takes a process-wide recursive lock, so
        # host threads driving different contexts (the app's --gpus mode) serialise call by call.
        entries = ("create|destroy|approx_table|sync|tile|set_tile_lists|get_lists|row_costs|render_device|frame_render|"
                   "fp32_peak|approx_rate|term_peak|mix_peak|set_host_pinning|pin_buffer|unpin_buffer")
        src, k = re.subn(r"^((?:int|void) vrt_cuda_(?:" + entries + r")\([^;{]*\)\n\{\n)", r"\1    emu::ApiLock emu_api_lock_;\n", src, flags=re.M)
        assert k == 17, f"{k} C-ABI entry definitions found for the interpreter's API lock (17 expected)"
        src, k = re.subn(r"^(static int (?:set_gaussians_impl|render_host)\([^;{]*\)\n\{\n)", r"\1    emu::ApiLock emu_api_lock_;\n", src, flags=re.M)
        assert k == 2
    assert "extern __shared__" not in src
    assert "<<<" not in src and not re.search(r"\basm\b", src), f"{name}: untranslated CUDA construct left"
    return src


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh") or f == "vrt_cuda.cu")


def build(force=False, asan=False):
    """Translate + compile if the product sources (or the interpreter) changed; returns the library path.
    asan=True: the same translation with -fsanitize=address into _build/asan/ (load it in a process started with
    LD_PRELOAD=libasan.so): out-of-bounds READS and writes of the kernels and the host code abort with the source line."""
    out_dir = os.path.join(BUILD, "asan") if asan else BUILD
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libvrt_cuda_emu.so")
    h = hashlib.sha256()
    inputs = [os.path.join(CSRC, f) for f in sources()] + [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "build_emu.py")]
    inputs += [os.path.join(ROOT, "include", f) for f in ("vrt_cuda.h", "vrt_approx_tables.h")]
    for p in inputs:
        h.update(open(p, "rb").read())
    stamp = os.path.join(out_dir, "stamp")
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return lib
    for f in sources():
        out = translate(f, open(os.path.join(CSRC, f)).read())
        dst = "vrt_cuda_emu.cpp" if f == "vrt_cuda.cu" else f
        with open(os.path.join(out_dir, dst), "w") as fh:
            fh.write(f"// GENERATED by tests/emu/build_emu.py from simd-gaussian-ray-tracing_b200/csrc/{f} -- test infrastructure, do not edit\n" + out)
    opt = ["-O1", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else ["-O2"]
    cmd = ["g++", "-std=c++17", *opt, "-g", "-march=x86-64-v3", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-unknown-pragmas", "-Wno-attributes",
           "-I", out_dir, "-I", HERE, "-I", os.path.join(ROOT, "include"),
           "-o", lib, os.path.join(out_dir, "vrt_cuda_emu.cpp"), os.path.join(HERE, "cuda_emu.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulated build failed:\n" + r.stdout[-6000:])
    with open(stamp, "w") as fh:
        fh.write(h.hexdigest())
    return lib


def libasan():
    """path of the AddressSanitizer runtime of this g++ (to LD_PRELOAD), or None"""
    r = subprocess.run(["g++", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True)
    p = r.stdout.strip()
    return p if r.returncode == 0 and os.path.isabs(p) and os.path.exists(p) else None


def build_selftest():
    """tests/emu/selftest.cu (kernels that pin the interpreter itself) -> _build/libemu_selftest.so"""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libemu_selftest.so")
    with open(os.path.join(BUILD, "selftest_emu.cpp"), "w") as fh:
        fh.write(translate("selftest.cu", open(os.path.join(HERE, "selftest.cu")).read()))
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes", "-I", HERE,
           "-o", out, os.path.join(BUILD, "selftest_emu.cpp"), os.path.join(HERE, "cuda_emu.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("interpreter self-test build failed:\n" + r.stdout[-6000:])
    return out


if __name__ == "__main__":
    print(build(force=True))
    print(build_selftest())
