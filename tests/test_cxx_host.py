"""The C++ host side above the C ABI (include/rbfmap_mapping.hpp: the mirror of precice::mapping::Mapping for the three RBF
mappings) and its parity test tests/cpp/test_mapping_cxx.cpp - the reference's own test scenarios and golden literals
driven from C++ exactly like the reference's unit tests drive a Mapping, linked against librbfmap.so only.

CPU (`-m "not gpu"`): the header and the test compile warning-free as C++17, link against the library, and the binary
reports the missing device (no compute without a GPU). GPU (`-m gpu`): the binary runs all scenarios."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_mapping_cxx.cpp")
LIBDIR = os.path.join(ROOT, "precice_b200")


def _build(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    if not os.path.exists(os.path.join(LIBDIR, "librbfmap.so")):
        import precice_b200  # noqa: F401  builds the library in-tree
        from precice_b200 import _lib
        _lib.build()
    exe = str(tmp_path / "test_mapping_cxx")
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lrbfmap", f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath-link,{cuda_lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


def _env():
    env = dict(os.environ)
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    env["LD_LIBRARY_PATH"] = cuda_lib + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


def test_cxx_host_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    import precice_b200
    if precice_b200.device_count() == 0:  # the product has no CPU fallback: the binary says so and stops
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=_env())
        assert r.returncode == 2 and "no CUDA device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cxx_host_reference_scenarios(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=_env())
    assert r.returncode == 0 and "CXX MAPPING OK" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
