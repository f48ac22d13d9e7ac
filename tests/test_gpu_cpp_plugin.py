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
