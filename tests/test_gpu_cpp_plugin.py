"""Runs the C++ tile plug-in test (include/tadev.hpp over the C ABI; tests/cpp/test_tile_plugin.cpp)
on the GPU: the reference-language side of the drop-in boundary."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_tile_plugin():
    exe = os.path.join(ROOT, "tests", "cpp", "build", "test_tile_plugin")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "tiledarray_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "CPP_PLUGIN OK" in out.stdout, out.stdout + out.stderr


def _run(exe_name, args=(), expect=None):
    exe = os.path.join(ROOT, *exe_name)
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "tiledarray_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and (expect is None or expect in out.stdout), out.stdout + out.stderr
    return out.stdout


def test_cpp_tiledarray_api():
    """include/tiledarray.hpp: TA::TArrayD / TSpArrayD, c("m,n") = a("m,k") * b("k,n") and friends in C++
    (tests/cpp/test_tiledarray_api.cpp restates the reference's own tests of the path)."""
    _run(("tests", "cpp", "build", "test_tiledarray_api"), expect="ALL TILEDARRAY API TESTS PASSED")


def test_cpp_ta_dense_example():
    """examples/ta_dense.cpp == the reference's examples/gemm/ta_dense.cpp on the device engine."""
    out = _run(("examples", "ta_dense"), ("2048", "256", "2"), expect="Verification        = passed")
    assert "Median GFLOPS" in out
