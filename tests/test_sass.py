"""Static checks on the built sm_100a code (cuobjdump, no GPU needed).

* The reprojection must stay bit-exact against the oracle: ptxas 12.9 contracts `mul.rn.f32x2` followed by
  `add/sub.rn.f32x2` into FFMA2 even with --fmad=false, so K1 keeps its products scalar and only packs sums
  (csrc/bmfr_device.cuh).  A fused multiply-add in the K1 kernels would silently break parity.
* The FUSED kernels are the TMA-fed kernels DESIGN.md describes: their SASS must show it (UTMALDG = cp.async.bulk.tensor,
  SYNCS = mbarrier, FFMA2 = packed fp32 fma).
"""
import re
import shutil
import subprocess

import pytest

from bmfr_b200 import _lib


def _sass_by_function():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    _lib.load()
    out = subprocess.run([exe, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", line):
            funcs[cur].append(line)
    return funcs


@pytest.fixture(scope="module")
def sass():
    if not (shutil.which("cuobjdump") or shutil.os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not available")
    return _sass_by_function()


def _ops(lines):
    return [re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?", "", l).split()[0].split(".")[0] for l in lines]


def test_reprojection_has_no_fused_multiply_add_pairs(sass):
    k1 = {n: l for n, l in sass.items() if "reproject_kernel" in n or "k1_accumulate_noisy_kernel" in n}
    assert len(k1) >= 4, list(sass)
    for name, lines in k1.items():
        ops = _ops(lines)
        assert "FFMA2" not in ops, f"{name}: a packed multiply-add was contracted — K1 is no longer bit-exact"
        assert "FADD2" in ops, f"{name}: the packed sums are gone"
        # scalar FFMA only inside the IEEE division / sqrt sequences (they come with MUFU.RCP + FCHK)
        assert ops.count("FFMA") <= 12 * ops.count("FCHK") + 8, f"{name}: unexpected scalar FFMA count {ops.count('FFMA')}"


@pytest.mark.parametrize("kernel,count,min_ffma2", [("fit_qr_kernel", 2, 300), ("fit_gram_kernel", 6, 80)])
def test_fit_kernels_use_tma_and_packed_fma(sass, kernel, count, min_ffma2):
    fit = [l for n, l in sass.items() if kernel in n]
    assert len(fit) == count, list(sass)  # whole image / strip (x the three feature lists for the Gram fit)
    for lines in fit:
        ops = _ops(lines)
        assert "UTMALDG" in ops, "the TMA tile loads of the fit are gone"
        assert ops.count("FFMA2") > min_ffma2, "the factorisation / the Gram products are no longer on packed pairs"
        assert "SYNCS" in ops, "mbarrier hand-off missing"
        assert len(lines) * 16 < 64 * 1024, f"fit kernel grew to {len(lines) * 16 // 1024} KB of SASS (I-cache)"
        # shared memory must be addressed as such: a struct reached through an integer-rounded pointer
        # turns every access into a generic LD / ST (what the kernel did until r01 v7)
        assert ops.count("LDS") > 100 and ops.count("LD") < 16, (ops.count("LDS"), ops.count("LD"))


def test_post_pass_stages_its_tiles_with_tma(sass):
    post = [l for n, l in sass.items() if "post_tma_kernel" in n]
    assert len(post) == 9, list(sass)  # (whole image, whole image with padded history, strip) x the three feature lists
    for lines in post:
        ops = _ops(lines)
        assert ops.count("UTMALDG") >= 6, "the six bulk tensor copies of a tile (normals, positions, albedo, prev_pixels, accept, spp)"
        assert "SYNCS" in ops, "mbarrier hand-off missing"
        assert ops.count("LDS") > 100 and ops.count("LD") < 16, (ops.count("LDS"), ops.count("LD"))


def test_strip_kernels_store_to_peers_in_wide_rows(sass):
    """The in-kernel halo exchange: the strip instantiations poll with volatile loads (LDG.E.STRONG / .SYS), fence at system
    scope and send staged rows as 128- / 64-bit stores."""
    for name, width in (("reproject_kernelILb1E", "128"), ("post_tma_kernelILb1E", "64")):
        lines = next(l for n, l in sass.items() if name in n)
        text = "\n".join(lines)
        assert "MEMBAR" in text and ".SYS" in text, f"{name}: no system-scope fence"
        assert re.search(r"STG\.E\.(\w+\.)*" + width, text), f"{name}: no {width}-bit global stores"


def test_post_pass_has_its_interior_path_and_the_shared_arithmetic(sass):
    """The interior tiles' phase A starts with a warp-uniform test per pixel pair (VOTE.ALL), the weighted sum clamps with the
    NaN-propagating maximum, the tone map multiplies with saturation; the padded-history instantiation gathers with
    128-bit loads."""
    for name, lines in sass.items():
        if "post_tma_kernel" not in name:
            continue
        text = "\n".join(lines)
        assert "VOTE.ALL" in text, f"{name}: the per-pair footprint test of the interior path is gone"
        assert re.search(r"FMNMX(3)?\.NAN", text), f"{name}: clamp_negative_nan no longer compiles to a NaN-propagating maximum"
        assert re.search(r"FMUL(\.\w+)*\.SAT", text), f"{name}: the tone map's saturating multiply is gone"
        if "ELi4ELi4E" in name:  # <.., PX = 4, HS = 4>
            assert len(re.findall(r"LDG\.E\.128", text)) >= 24, f"{name}: padded history is not gathered with 128-bit loads"


def test_fit_draws_its_next_block_ahead_of_the_barrier(sass):
    """fit_gram_kernel: the atomic on the block counter is issued before the block barrier that precedes its use (the L2
    round trip used to sit behind it on thread 0's path)."""
    for name, lines in sass.items():
        if "fit_gram_kernel" not in name:
            continue
        ops = _ops(lines)
        atom = next(i for i, o in enumerate(ops) if o.startswith("ATOMG") or o.startswith("ATOM"))
        bars = [i for i, o in enumerate(ops) if o == "BAR"]
        assert any(b > atom for b in bars), name
        # between the draw and the next CTA barrier lies the whole first phase of a block: its tile loads
        nxt = min(b for b in bars if b > atom)
        assert ops[atom:nxt].count("LDS") >= 72, (name, ops[atom:nxt].count("LDS"))


def _region_counts(tmp_path, source, kernel, regions):
    """scripts/sass_region_count.py on the built library: static SASS count per source-line range, following inline chains."""
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    subprocess.run([exe, "-xelf", source, str(_lib.LIB_PATH)], cwd=tmp_path, check=True, capture_output=True)
    cubin = next(tmp_path.glob("*.cubin"))
    args = [f"{f}:{lo}-{hi}:{name}" for name, (f, lo, hi) in regions.items()]
    out = subprocess.run([sys.executable, str(root / "scripts" / "sass_region_count.py"), str(cubin), kernel, *args], capture_output=True,
                         text=True, check=True).stdout
    return {l.split()[1].rstrip(":"): int(l.split()[0]) for l in out.splitlines() if l.strip()}


def _line_of(path, marker):
    for i, l in enumerate(path.read_text().splitlines(), 1):
        if marker in l:
            return i
    raise AssertionError(f"{marker!r} not found in {path}")


def test_fit_interior_blocks_load_their_tile_at_immediate_offsets(tmp_path):
    """The tile loads of a block inside the image: 72 LDS from one base address; the mirrored path next to it pays per-row index
    arithmetic and branches (what the source-level profile showed at 17 % of the kernel's instructions)."""
    from pathlib import Path
    if not shutil.which("nvdisasm") and not Path("/usr/local/cuda/bin/nvdisasm").exists():
        pytest.skip("nvdisasm not available")
    src = Path(__file__).resolve().parent.parent / "bmfr_b200" / "csrc" / "bmfr_fit.cu"
    a = _line_of(src, "if (BMFR_GRAM_INTERIOR_LOADS && by_tma")
    b = _line_of(src, "} else if (by_tma) {")
    counts = _region_counts(tmp_path, "bmfr_fit", "fit_gram_kernelILb0ELi0",
                            {"interior": ("bmfr_fit.cu", a, b - 1), "mirrored": ("bmfr_fit.cu", b, b + 13)})
    assert 72 <= counts["interior"] <= 110, counts
    assert counts["mirrored"] >= 2 * counts["interior"], counts
