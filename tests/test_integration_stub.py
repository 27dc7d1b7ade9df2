"""INTEGRATION.md shows the lines a maintainer of the reference would put in place of bmfr.cpp:183-249,315-517.
That stub must stay valid against include/bmfr_b200.h: it is extracted here, given the declarations it relies on in
bmfr.cpp (the dataset header's arrays, the per-frame vectors of bmfr.cpp:253-258) and compiled."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

PRELUDE = r"""
#include <cstdio>
#include <vector>
#define IMAGE_WIDTH 1280
#define IMAGE_HEIGHT 720
#define FRAME_COUNT 60
// what bmfr.cpp has in scope at this point: camera_matrices.h (bmfr.cpp:47) and the loaded frames (bmfr.cpp:253-258)
extern const float position_limit_squared, normal_limit_squared;
extern const float camera_matrices[FRAME_COUNT][4][4];
extern const float pixel_offsets[FRAME_COUNT][2];
extern std::vector<float> albedos[FRAME_COUNT], normals[FRAME_COUNT], positions[FRAME_COUNT], noisy_input[FRAME_COUNT],
    out_data[FRAME_COUNT];
"""


def test_the_stub_of_integration_md_compiles_against_the_header(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    text = (ROOT / "INTEGRATION.md").read_text()
    m = re.search(r"## The stub.*?```cpp\n(.*?)```", text, re.S)
    assert m, "INTEGRATION.md lost its stub"
    code = m.group(1)
    include, body = code.split("\n", 1)
    assert include.strip() == '#include "bmfr_b200.h"'
    body = body.replace("        ...\n", "        (void)ms;\n")
    src = tmp_path / "stub.cpp"
    src.write_text(include + "\n" + PRELUDE + "int tasks() {\n" + body + "    return 0;\n}\n")
    r = subprocess.run([gxx, "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT / 'include'}", str(src)], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
