"""CPU-side checks of the product library: it loads without a GPU, exports every symbol
include/hnsw_b200.h declares, keeps the reference's struct layouts, and FAILS LOUDLY (no CPU fallback)
when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hnsw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-zA-Z_][a-zA-Z0-9_]*)\s*\(", src)
    pat = re.compile(r"^(init_|new_hnsw|drop_hnsw|insert_|parallel_|search_neighbours|file_dump|load_hnsw|get_hnswio|hnsw_b200_)")
    out = []
    for n in names:
        if not pat.match(n) or n.startswith("hnsw_b200_filter_fn"):
            continue
        if n not in out:
            out.append(n)
    return out


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.load_library()
    syms = declared_symbols()
    assert "parallel_search_neighbours_f32" in syms and "parallel_insert_f32" in syms and len(syms) > 90
    for s in syms:
        assert hasattr(L, s), f"libhnsw_b200.so does not export {s}"


def test_reference_struct_layouts(pkg):
    from importlib import import_module
    h = import_module("hnswlib-rs_b200.hnsw")
    # libext.rs:58-71,82-87 on x86-64
    assert C.sizeof(h.Neighbour_api) == 16 and h.Neighbour_api.d.offset == 8
    assert C.sizeof(h.Neighbourhood_api) == 16 and h.Neighbourhood_api.neighbours.offset == 8
    assert C.sizeof(h.Vec_api) == 16 and h.Vec_api.ptr.offset == 8


def test_no_gpu_means_loud_failure_not_fallback(pkg):
    L = pkg.load_library()
    if L.hnsw_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.HnswError) as e:
        pkg.Hnsw(16, 1000, 16, 200, "DistL2")
    assert "CUDA" in str(e.value) and "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    """the product path must not import / link / call anything under oracle/"""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "hnswlib-rs_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"pyoracle|liboracle|oracle/|import oracle", txt):
                    # comments that only NAME the mirrored file are allowed
                    for line in txt.splitlines():
                        if re.search(r"pyoracle|liboracle|import oracle", line) or (
                                "oracle/" in line and not line.lstrip().startswith(("//", "#", "*"))):
                            bad.append((f, line.strip()))
    assert not bad, bad
