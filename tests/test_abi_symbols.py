"""The C-ABI library loads and exports every symbol include/b2v.h declares (CPU-only:
no compute call is made)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    txt = (ROOT / "include" / "b2v.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2v_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    from invesalius3_b200 import _build, _lib
    _build.build_cuda()
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"libb2v.so does not export {s}"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes and include/b2v.h disagree"
    assert lib.b2v_version() >= 100
    assert lib.b2v_last_error() is not None


def test_sass_is_sm100a():
    import subprocess
    from invesalius3_b200 import _build
    lib = _build.build_cuda()
    out = subprocess.run(["cuobjdump", "-lelf", str(lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    for p in (ROOT / "invesalius3_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f"{p} imports the oracle"
    for p in (ROOT / "invesalius3_b200" / "csrc").glob("*"):
        if p.is_file() and p.suffix in (".cu", ".cuh", ".h"):
            assert "oracle" not in p.read_text().lower() or p.name == "mc_tables.h", p
