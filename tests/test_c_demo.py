"""The C ABI from plain C: examples/c_demo.c includes include/hnsw_b200.h as strict C99, links libhnsw_b200.so and calls
the reference's entry points (init_hnsw_f32, parallel_insert_f32, search_neighbours_f32, parallel_search_neighbours_f32)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, pkg):
    pkg.load_library()
    exe = str(tmp_path / "c_demo")
    lib = os.path.join(ROOT, "hnswlib-rs_b200", "lib")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_demo.c"), "-L" + lib, "-lhnsw_b200", "-Wl,-rpath," + lib, "-o", exe])
    return exe


def test_header_is_plain_c_and_fails_loudly_without_gpu(tmp_path, pkg):
    exe = _build(tmp_path, pkg)
    if pkg.load_library().hnsw_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_demo_runs_on_gpu(tmp_path, pkg):
    exe = _build(tmp_path, pkg)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "c_demo ok" in r.stdout, r.stdout + r.stderr
