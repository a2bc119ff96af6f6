"""CPU checks of oracle/cl_device.py (the OpenCL host that runs the reference's own program on the GPU box):
what can be verified without an OpenCL driver -- the embedded program text, the portability rewrite, the ICD
dispatch-table slots and the kernel list against the reference's kernel_type.go."""
import re

import pytest

from oracle import ref_binding

pytestmark = pytest.mark.skipif(not ref_binding.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def test_embedded_program_and_portability_rewrite():
    from oracle import cl_device as D

    raw, port = D.program_source(portable=False), D.program_source()
    kernels = set(re.findall(rb"__kernel\s+void\s+(\w+)", raw))
    assert set(k.encode() for k in D._KERNELS) <= kernels  # every kernel the host launches exists in the program
    assert len(kernels) == 17 == len(D._KERNELS)  # 11 on the path + 6 debug (tracer/opencl/kernel_type.go)
    assert b"#include" not in raw  # include-expanded: the GPU box has no /root/reference to resolve them against
    # exactly the 14 functional casts change, line count and everything else stay
    a, b = raw.split(b"\n"), port.split(b"\n")
    assert len(a) == len(b)
    changed = [(x, y) for x, y in zip(a, b) if x != y]
    assert len(changed) == 14
    for x, y in changed:
        assert re.sub(rb"(?<![\w)])\b(uint|int|uchar)\(", lambda m: b"(" + m.group(1) + b")(", x) == y
    assert not D._FUNCTIONAL_CAST.search(port)
    # casts that already were C casts and identifiers ending in a type name are left alone
    assert D._FUNCTIONAL_CAST.sub(b"X", b"(uint)(x) + print(1) + myint(2) + (float)(3)") == b"(uint)(x) + print(1) + myint(2) + (float)(3)"


def test_dispatch_slots_follow_cl_khr_icd_order():
    from oracle import cl_device as D

    names = [D._DISPATCH[k][0] for k in sorted(D._DISPATCH)]
    # the OpenCL 1.0 block of the ICD dispatch table is in cl.h declaration order
    assert names == ["clGetPlatformIDs", "clGetPlatformInfo", "clGetDeviceIDs", "clGetDeviceInfo", "clCreateContext",
                     "clReleaseContext", "clCreateCommandQueue", "clReleaseCommandQueue", "clCreateBuffer", "clReleaseMemObject",
                     "clCreateProgramWithSource", "clReleaseProgram", "clBuildProgram", "clGetProgramBuildInfo", "clCreateKernel",
                     "clReleaseKernel", "clSetKernelArg", "clFinish", "clEnqueueReadBuffer", "clEnqueueWriteBuffer",
                     "clEnqueueNDRangeKernel"]
    assert D._DISPATCH[59][0] == "clEnqueueNDRangeKernel" and D._DISPATCH[47][0] == "clFinish"


def test_unavailable_without_a_driver_is_reported_not_raised():
    from oracle import cl_device as D

    assert D.available() in (True, False)
    import bench

    if not D.available():
        r = bench.opencl_reference_sample(None, 8, 8, 1, 2)
        assert "unavailable" in r
