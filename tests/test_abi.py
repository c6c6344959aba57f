"""CPU: the C-ABI library builds, loads and exports exactly what include/sibgpu.h declares; without a GPU every
compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import sibelia_b200 as sb
from sibelia_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sibgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sibgpu_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree(built):
    assert header_symbols() == sorted(binding.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    lib = sb.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.sibgpu_version().startswith(b"sibgpu")


def test_no_cpu_fallback(built):
    """On a box without CUDA the product refuses to run instead of silently computing on the host."""
    if sb.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(sb.SibgpuError) as e:
        sb.Context(0)
    assert e.value.status == 1


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under sibelia_b200/ may import, link, load or execute it."""
    pkg = os.path.join(ROOT, "sibelia_b200")
    banned = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|liboracle|libsibelia_ref|_ref/|ref_shim|enum_restate)")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not banned.search(text), os.path.join(dirpath, f)
