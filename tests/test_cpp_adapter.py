"""The header-only C++ host layer (include/cbird_b200.hpp) — the shape a cbird `Index` subclass forwards
to — compiled with plain g++ against libcbird_b200.so. CPU: soft errors without a device; GPU: real run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def adapter_exe(cb, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "adapter_test")
    libdir = os.path.join(ROOT, "cbird_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp"),
           "-o", exe, "-L", libdir, "-lcbird_b200", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)
    return exe


def test_adapter_without_device(adapter_exe):
    import ctypes

    import cbird_b200

    n = ctypes.c_int(0)
    cbird_b200.lib().cb_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    out = subprocess.run([adapter_exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("NO_DEVICE OK")


@pytest.mark.gpu
def test_adapter_on_gpu(adapter_exe):
    out = subprocess.run([adapter_exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK ")
