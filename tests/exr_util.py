"""An independent OpenEXR scanline writer for the tests (numpy + zlib, following the OpenEXR file layout):
single part, compression NONE / RLE / ZIPS / ZIP, HALF and FLOAT channels."""
import struct
import zlib

import numpy as np

HALF, FLOAT = 1, 2
NONE, RLE, ZIPS, ZIP, PIZ = 0, 1, 2, 3, 4


def _attr(name, typ, payload):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def _predict(raw: bytes) -> bytes:
    """OpenEXR's byte transform before zlib / RLE: split even / odd bytes, then delta-encode."""
    a = np.frombuffer(raw, dtype=np.uint8)
    t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)
    d = t.copy()
    d[1:] = (t[1:] - t[:-1] + 128 + 256) & 0xFF
    return d.astype(np.uint8).tobytes()


def _rle(data: bytes) -> bytes:
    out, i, n = bytearray(), 0, len(data)
    while i < n:
        j = i
        while j + 1 < n and data[j + 1] == data[i] and j - i < 126:
            j += 1
        if j - i >= 2:  # a run of j - i + 1 equal bytes
            out += struct.pack("b", j - i) + data[i:i + 1]
            i = j + 1
        else:  # literals up to the next run of three
            j = i
            while j < n and j - i < 127 and not (j + 2 < n and data[j] == data[j + 1] == data[j + 2]):
                j += 1
            out += struct.pack("b", -(j - i)) + data[i:j]
            i = j
    return bytes(out)


def write_exr(path, planes, types, compression, xmin=0, ymin=0, decreasing=False, version=2, flags=0):
    """planes: {channel name: float32 array [H, W]}; types: {name: HALF | FLOAT}."""
    names = sorted(planes)  # the channel list of an EXR file is alphabetical
    h, w = next(iter(planes.values())).shape
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", types[n], 0, 1, 1) for n in names) + b"\0"
    box = struct.pack("<4i", xmin, ymin, xmin + w - 1, ymin + h - 1)
    header = struct.pack("<ii", 20000630, version | flags)
    header += _attr("channels", "chlist", chlist) + _attr("compression", "compression", bytes([compression]))
    header += _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box)
    header += _attr("lineOrder", "lineOrder", bytes([1 if decreasing else 0]))
    header += _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    header += _attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0)) + _attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    header += b"\0"
    lines = 16 if compression in (ZIP,) else 1
    chunks = []
    for y0 in range(0, h, lines):
        raw = b""
        for y in range(y0, min(y0 + lines, h)):
            for n in names:
                row = planes[n][y]
                raw += row.astype("<f2").tobytes() if types[n] == HALF else row.astype("<f4").tobytes()
        if compression == NONE:
            data = raw
        else:
            data = zlib.compress(_predict(raw)) if compression in (ZIPS, ZIP) else _rle(_predict(raw))
            if len(data) >= len(raw):
                data = raw  # OpenEXR stores the chunk raw when compression does not help
        chunks.append((y0 + ymin, data))
    order = list(reversed(range(len(chunks)))) if decreasing else list(range(len(chunks)))
    table_at = len(header)
    pos = table_at + 8 * len(chunks)
    offsets, body = [0] * len(chunks), b""
    for k in order:  # the offset table is in increasing y, the chunks themselves follow the line order
        offsets[k] = pos + len(body)
        body += struct.pack("<ii", chunks[k][0], len(chunks[k][1])) + chunks[k][1]
    with open(path, "wb") as f:
        f.write(header + struct.pack(f"<{len(chunks)}Q", *offsets) + body)




# ---- the image of tests/golden/exr/*.exr (written by tests/golden/make_exr_fixtures.py with the OpenEXR library) ----
FIXTURE_W, FIXTURE_H, FIXTURE_SEED = 24, 19, 20261018
FIXTURE_TALL_W, FIXTURE_TALL_H = 21, 70


def fixture_image(tall=False):
    w, h = (FIXTURE_TALL_W, FIXTURE_TALL_H) if tall else (FIXTURE_W, FIXTURE_H)
    img = np.random.default_rng(FIXTURE_SEED + int(tall)).normal(0, 2, (h, w, 3)).astype(np.float32)
    img[: h // 2] = np.linspace(0, 4, w, dtype=np.float32)[None, :, None] * np.array([1, -1, 0.5], dtype=np.float32)
    if tall:
        img[h // 2: h // 2 + 8] = 0.75  # constant rows: the Huffman run-length symbol
    return img
