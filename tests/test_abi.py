"""CPU tests: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a GPU (no CPU fallback)."""
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def header_symbols():
    # the drop-in header and the bench-only header (edmd_cuda_bench.h: not part of the interface)
    text = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(edmd_cuda_\w+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert header_symbols() == sorted(pkg.binding.SYMBOLS)


def test_library_exports_every_header_symbol(pkg):
    lib = pkg.load_library()
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_box_struct_layout_matches_header(pkg):
    # 3 x int32 (+4 pad) + 9 doubles
    import ctypes
    assert ctypes.sizeof(pkg.binding.Box) == 16 + 9 * 8


def test_product_never_imports_oracle():
    for p in list((ROOT / "graphical-edmd_b200").rglob("*.py")) + \
            list((ROOT / "graphical-edmd_b200").rglob("*.c*")) + \
            list((ROOT / "graphical-edmd_b200").rglob("*.h")):
        txt = p.read_text(errors="ignore")
        assert "oracle" not in txt.lower() or p.name == "__init__.py", p


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_create_fails_loudly_without_gpu(pkg):
    with pytest.raises(pkg.EdmdError):
        pkg.EdmdCuda(100, 30.0, 30.0)


def test_synth_is_deterministic_and_non_overlapping(pkg):
    a = pkg.synth.lattice_config(3000, 0.8, seed=9, small_fraction=0.3)
    b = pkg.synth.lattice_config(3000, 0.8, seed=9, small_fraction=0.3)
    for k in ("x", "y", "vx", "vy", "rad"):
        assert np.array_equal(a[k], b[k])
    n, lx, ly = a["n"], a["lx"], a["ly"]
    assert abs(a["vx"].mean()) < 1e-12 and abs(a["vy"].mean()) < 1e-12
    assert (a["x"] >= 0).all() and (a["x"] < lx).all() and (a["y"] >= 0).all() and (a["y"] < ly).all()
    dx = a["x"][:, None] - a["x"][None, :]
    dy = a["y"][:, None] - a["y"][None, :]
    dx -= lx * np.rint(dx / lx)
    dy -= ly * np.rint(dy / ly)
    d2 = dx * dx + dy * dy + 1e9 * np.eye(n)
    assert (d2 >= 4 * a["rad"][:, None] * a["rad"][None, :]).all()


def test_bench_helper_is_not_in_the_drop_in_header():
    assert "edmd_cuda_bench" not in (ROOT / "include" / "edmd_cuda.h").read_text()
