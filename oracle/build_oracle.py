"""TEST INFRASTRUCTURE — builds the two CPU checkers (see oracle/bmfr_oracle.h).

  port      : oracle/bmfr_oracle.c -> oracle/libbmfr_oracle.so               (always)
  reference : /root/reference/opencl/bmfr.cl, compiled as C++ through oracle/cl_shim/,
              -> oracle/_ref/libbmfr_clref.so   (only where /root/reference exists; the GPU box
              uses the prebuilt file that travels with the repo snapshot)

The reference's sources are read where they lie; nothing of them is written outside oracle/_ref/,
and the intermediate files are removed after the compile.
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REFERENCE = Path(os.environ.get("BMFR_REFERENCE_DIR", "/root/reference")) / "opencl"
PORT_LIB = HERE / "libbmfr_oracle.so"
REF_DIR = HERE / "_ref"
REF_LIB = REF_DIR / "libbmfr_clref.so"

CFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared"]

# Feature lists (include/bmfr_b200.h, bmfr_feature_set).  Set 0 is read from bmfr.cpp itself; the others are what a user
# of the reference would write into NOT_SCALED_FEATURE_BUFFERS / SCALED_FEATURE_BUFFERS (bmfr.cpp:63-77).
FEATURE_SETS = {
    1: ("1.f,normal.x,normal.y,normal.z,", "world_position.x,world_position.y,world_position.z"),
    2: ("1.f,", "world_position.x,world_position.y,world_position.z,world_position.x*world_position.x,"
                "world_position.y*world_position.y,world_position.z*world_position.z"),
}


def _suffix(feature_set: int) -> str:
    return "" if feature_set == 0 else f"_fs{feature_set}"


def port_lib(feature_set=0) -> Path:
    return HERE / f"libbmfr_oracle{_suffix(feature_set)}.so"


# The reference's two fitter tuning toggles (bmfr.cpp:82-84 -> bmfr.cl:100-119,609-649,664-688): builds of the shim library with
# the other value, for tests/test_oracle_pin.py (test_compressed_r_toggle_*, test_cache_tmp_data_toggle_*).
TOGGLE_VARIANTS = {"r0": {"COMPRESSED_R": 0}, "c0": {"CACHE_TMP_DATA": 0}}


def ref_lib(feature_set=0, variant="") -> Path:
    return REF_DIR / f"libbmfr_clref{_suffix(feature_set)}{'_' + variant if variant else ''}.so"


def clgpu_lib(feature_set=0) -> Path:
    return REF_DIR / f"libbmfr_clgpu{_suffix(feature_set)}.so"

# OpenCL C constructs that C++ cannot express through the shim header alone.  Each rewrite is purely
# syntactic, must match exactly `count` times, and is listed in DESIGN.md.
REWRITES = [
    # work-group-local variable declarations (as opposed to __local pointer qualifiers)
    (r"__local float u_length_squared, dot, block_min, block_max, vec_length;",
     "CLSHIM_WG_LOCAL float u_length_squared, dot, block_min, block_max, vec_length;", 1),
    (r"local float3 divider;", "CLSHIM_WG_LOCAL float3 divider;", 1),
    # the one vector ternary of the file (C++ cannot overload ?:)
    (r"color = color < 0.f ? 0.f : color;", "color = vec_ternary(color < 0.f, 0.f, color);", 1),
]


def _stamp(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(Path(p).read_bytes())
    return h.hexdigest()


def build_port(force=False, feature_set=0) -> Path:
    src = [HERE / "bmfr_oracle.c", HERE / "bmfr_oracle.h"]
    lib = port_lib(feature_set)
    stamp = HERE / "_build" / f"port{_suffix(feature_set)}.sha256"
    digest = _stamp(src, " ".join(CFLAGS) + str(feature_set))
    if not force and lib.exists() and stamp.exists() and stamp.read_text() == digest:
        return lib
    stamp.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-std=gnu11", *CFLAGS, f"-DBMFR_FEATURE_SET={feature_set}", "-o", str(lib), str(src[0]), "-lm"]
    subprocess.run(cmd, check=True)
    stamp.write_text(digest)
    return lib


def _defines_from_bmfr_cpp(text: str, feature_set=0, toggles=None) -> str:
    """The -D options bmfr.cpp:205-232 derives from its own #defines (bmfr.cpp:56-118); feature_set != 0 replaces the two
    feature strings the way an edit of bmfr.cpp:65-77 would."""
    def strings_of(name):
        m = re.search(r"#define %s \\\n((?:\".*\"\\?\n)+)" % name, text)
        if not m:
            raise RuntimeError(f"{name} not found in bmfr.cpp")
        return "".join(re.findall(r"\"(.*?)\"", m.group(1)))
    not_scaled, scaled = strings_of("NOT_SCALED_FEATURE_BUFFERS"), strings_of("SCALED_FEATURE_BUFFERS")
    if feature_set != 0:
        not_scaled, scaled = FEATURE_SETS[feature_set]
    n_ns = not_scaled.count(",")          # bmfr.cpp:195-196
    n_s = scaled.count(",") + 1           # bmfr.cpp:198-199
    buffers = n_ns + n_s + 3              # bmfr.cpp:202
    out = [
        "// generated from /root/reference/opencl/bmfr.cpp by oracle/build_oracle.py — do not commit",
        f"#define BUFFER_COUNT {buffers}",
        f"#define FEATURES_NOT_SCALED {n_ns}",
        f"#define FEATURES_SCALED {n_s}",
        f"#define FEATURE_BUFFERS {not_scaled}{scaled}",
        f"#define R_EDGE {buffers - 2}",
    ]
    for name in ("LOCAL_WIDTH", "LOCAL_HEIGHT", "BLOCK_EDGE_LENGTH", "LOCAL_SIZE", "COMPRESSED_R", "CACHE_TMP_DATA",
                 "ADD_REQD_WG_SIZE"):
        m = re.search(r"^#define %s (\d+)\s*$" % name, text, re.M)
        if not m:
            raise RuntimeError(f"{name} not found in bmfr.cpp")
        out.append(f"#define {name} {(toggles or {}).get(name, m.group(1))}")
    out.append("#define BLOCK_PIXELS (BLOCK_EDGE_LENGTH * BLOCK_EDGE_LENGTH)")
    return "\n".join(out) + "\n"


def build_reference(force=False, feature_set=0, variant=""):
    """Returns the path of the reference-kernel library, or None when it cannot be (re)built.  variant: a key of
    TOGGLE_VARIANTS (the reference's fitter tuning toggles set the other way), "" = bmfr.cpp as shipped."""
    REF_LIB, cl, cpp = ref_lib(feature_set, variant), REFERENCE / "bmfr.cl", REFERENCE / "bmfr.cpp"
    toggles = TOGGLE_VARIANTS[variant] if variant else None
    shim = [HERE / "cl_shim" / "cl_shim.hpp", HERE / "cl_shim" / "cl_host.cpp", HERE / "bmfr_oracle.h"]
    if not cl.exists() or not cpp.exists():
        return REF_LIB if REF_LIB.exists() else None   # GPU box: use what travelled
    digest = _stamp([cl, cpp, *shim, Path(__file__)], " ".join(CFLAGS) + str(feature_set) + variant)
    stamp = REF_DIR / f"ref{_suffix(feature_set)}{variant}.sha256"
    if not force and REF_LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return REF_LIB
    work = REF_DIR / f"gen{_suffix(feature_set)}{variant}"
    work.mkdir(parents=True, exist_ok=True)
    try:
        src = cl.read_text()
        for old, new, count in REWRITES:
            if src.count(old) != count:
                raise RuntimeError(f"rewrite {old!r}: expected {count} match(es), found {src.count(old)}")
            src = src.replace(old, new)
        (work / "bmfr_cl.gen.inc").write_text(src)
        (work / "bmfr_defines.gen.h").write_text(_defines_from_bmfr_cpp(cpp.read_text(), feature_set, toggles))
        cmd = ["g++", "-std=gnu++17", *CFLAGS, "-Wno-narrowing", "-Wno-attributes", "-I", str(work),
               "-o", str(REF_LIB), str(HERE / "cl_shim" / "cl_host.cpp"), "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("g++ failed on the shim build of bmfr.cl")
        stamp.write_text(digest)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return REF_LIB


CLGPU_LIB = REF_DIR / "libbmfr_clgpu.so"


def _opencl_header(cl_text: bytes, cpp_text: str, feature_set=0) -> str:
    """bmfr.cl as a byte array + the feature-list build options of bmfr.cpp:63-77,195-202 — the unmodified
    kernel source is what the OpenCL driver on the GPU box compiles."""
    defs = dict(re.findall(r"^#define (\w+) (.*)$", _defines_from_bmfr_cpp(cpp_text, feature_set), re.M))
    out = ["// generated from /root/reference/opencl/{bmfr.cl,bmfr.cpp} by oracle/build_oracle.py — do not commit",
           f"#define BMFR_CL_FEATURE_OPTIONS \"{defs['FEATURE_BUFFERS']}\""]
    for name in ("BUFFER_COUNT", "FEATURES_NOT_SCALED", "FEATURES_SCALED", "LOCAL_WIDTH", "LOCAL_HEIGHT", "LOCAL_SIZE",
                 "COMPRESSED_R", "CACHE_TMP_DATA", "ADD_REQD_WG_SIZE"):
        out.append(f"#define BMFR_CL_{name} {defs[name]}")
    out.append(f"static const unsigned long kBmfrClSourceLen = {len(cl_text)};")
    out.append("static const unsigned char kBmfrClSource[] = {")
    for i in range(0, len(cl_text), 32):
        out.append(",".join(str(b) for b in cl_text[i:i + 32]) + ",")
    out.append("0};")
    return "\n".join(out) + "\n"


def build_opencl_host(force=False, feature_set=0):
    """oracle/_ref/libbmfr_clgpu.so: the reference's unmodified bmfr.cl + a host that drives it through the box's
    OpenCL ICD (oracle/cl_gpu/cl_gpu_host.c).  Needs no OpenCL at build time (the API is resolved with dlopen)."""
    CLGPU_LIB, cl, cpp = clgpu_lib(feature_set), REFERENCE / "bmfr.cl", REFERENCE / "bmfr.cpp"
    host = HERE / "cl_gpu" / "cl_gpu_host.c"
    if not cl.exists() or not cpp.exists():
        return CLGPU_LIB if CLGPU_LIB.exists() else None   # GPU box: use what travelled
    digest = _stamp([cl, cpp, host, HERE / "bmfr_oracle.h", Path(__file__)], str(feature_set))
    stamp = REF_DIR / f"clgpu{_suffix(feature_set)}.sha256"
    if not force and CLGPU_LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return CLGPU_LIB
    work = REF_DIR / f"gen_clgpu{_suffix(feature_set)}"
    work.mkdir(parents=True, exist_ok=True)
    try:
        (work / "bmfr_cl_source.gen.h").write_text(_opencl_header(cl.read_bytes(), cpp.read_text(), feature_set))
        cmd = ["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-Wall", "-I", str(work), "-o", str(CLGPU_LIB), str(host), "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed on the OpenCL host")
        stamp.write_text(digest)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return CLGPU_LIB


def build_all(force=False):
    out = []
    for fs in (0, *FEATURE_SETS):
        out.append((fs, build_port(force, fs), build_reference(force, fs), build_opencl_host(force, fs)))
    return out


if __name__ == "__main__":
    for fs, port, ref, clgpu in build_all("--force" in sys.argv):
        print(f"feature set {fs}: port {port}, reference {ref}, opencl {clgpu}")
