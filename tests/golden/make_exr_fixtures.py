#!/usr/bin/env python
"""Writes tests/golden/exr/*.exr with OpenCV, i.e. with the OpenEXR library itself (cv2 4.13 bundles it), so
that tests/test_io.py can pin the EXR reader against files it did not have a hand in.  The pixel values are
tests/exr_util.py:fixture_image(); the test regenerates them.

    OPENCV_IO_ENABLE_OPENEXR=1 python tests/golden/make_exr_fixtures.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from tests.exr_util import fixture_image  # noqa: E402

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
import cv2  # noqa: E402
import numpy as np  # noqa: E402

VARIANTS = {
    "none_f32": (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_NO),
    "rle_f16": (cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_RLE),
    "zips_f16": (cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_ZIPS),
    "zip_f32": (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_ZIP),
    "piz_f32": (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_PIZ),
    "piz_f16": (cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_PIZ),
    "pxr24_f32": (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_PXR24),  # not covered: must be refused
}
# a second, taller image for the 32-scanline PIZ chunks: three chunks, the last one ragged
TALL = {
    "piz_f32_tall": (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_PIZ),
    "zip_f16_tall": (cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_ZIP),
}

if __name__ == "__main__":
    out = Path(__file__).resolve().parent / "exr"
    out.mkdir(exist_ok=True)
    for variants, img in ((VARIANTS, fixture_image()), (TALL, fixture_image(tall=True))):
        bgr = np.ascontiguousarray(img[..., ::-1])  # OpenCV's channel order
        for name, (typ, comp) in variants.items():
            assert cv2.imwrite(str(out / f"{name}.exr"), bgr, [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, comp])
            print(name, (out / f"{name}.exr").stat().st_size, "bytes")
