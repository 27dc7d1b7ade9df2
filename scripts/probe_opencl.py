"""GPU-box probe: is NVIDIA's OpenCL ICD usable?  Prints what it finds."""
import ctypes as C
import glob
import os

print("vendors dir:", os.path.isdir("/etc/OpenCL/vendors"), glob.glob("/etc/OpenCL/vendors/*"))
cands = glob.glob("/usr/lib/libnvidia-opencl.so*") + glob.glob("/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so*") + \
    glob.glob("/usr/local/nvidia/lib*/libnvidia-opencl.so*")
print("icd libs:", cands)
for path in cands[:1]:
    try:
        lib = C.CDLL(path)
    except OSError as e:
        print("dlopen failed:", e)
        continue
    for sym in ("clGetPlatformIDs", "clIcdGetPlatformIDsKHR", "clGetExtensionFunctionAddress", "clGetPlatformInfo",
                "clGetDeviceIDs", "clCreateContext", "clBuildProgram", "clEnqueueNDRangeKernel"):
        print(f"  {sym}: {'exported' if hasattr(lib, sym) else 'missing'}")
    fn = getattr(lib, "clGetPlatformIDs", None) or getattr(lib, "clIcdGetPlatformIDsKHR", None)
    if fn:
        n = C.c_uint(0)
        plats = (C.c_void_p * 4)()
        rc = fn(4, plats, C.byref(n))
        print("  platforms:", rc, n.value)
        if rc == 0 and n.value:
            buf = C.create_string_buffer(256)
            sz = C.c_size_t()
            for what, code in (("name", 0x0902), ("version", 0x0901)):
                lib.clGetPlatformInfo(C.c_void_p(plats[0]), code, 256, buf, C.byref(sz))
                print(f"  platform {what}:", buf.value.decode())
            devs = (C.c_void_p * 8)()
            nd = C.c_uint(0)
            if hasattr(lib, "clGetDeviceIDs"):
                rc = lib.clGetDeviceIDs(C.c_void_p(plats[0]), C.c_ulonglong(4), 8, devs, C.byref(nd))   # CL_DEVICE_TYPE_GPU
                print("  gpu devices:", rc, nd.value)
try:
    ld = C.CDLL("/usr/local/cuda/lib64/libOpenCL.so.1")
    n = C.c_uint(0)
    rc = ld.clGetPlatformIDs(0, None, C.byref(n))
    print("ICD loader clGetPlatformIDs:", rc, n.value)
except OSError as e:
    print("no ICD loader:", e)
