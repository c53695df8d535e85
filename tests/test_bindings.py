"""The reference-side bindings of INTEGRATION.md: the C++ facade (include/SimpleCABAC.hpp) and the
MATLAB mexFunction shim (isscabac_b200/mex/SimpleCABACMex_b200.cpp, compiled against the stub mex.h
that stands in for MATLAB).  CPU: they compile and link against libisscabac.so.  GPU: they run and
reproduce the reference's known-answer vectors."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "isscabac_b200")
OUT = os.path.join(ROOT, "tests", "bindings", "_build")


def build(name, sources, extra_inc=()):
    import isscabac_b200 as I
    I.lib()   # makes sure libisscabac.so exists
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name)
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include")]
    for i in extra_inc:
        cmd += ["-I", i]
    cmd += sources + ["-L", LIBDIR, "-lisscabac", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def facade_exe():
    return build("facade_demo", [os.path.join(ROOT, "tests", "bindings", "facade_demo.cpp")])


def shim_exe():
    return build("mex_shim_driver",
                 [os.path.join(ROOT, "tests", "bindings", "mex_shim_driver.cpp"),
                  os.path.join(ROOT, "isscabac_b200", "mex", "SimpleCABACMex_b200.cpp")],
                 extra_inc=[os.path.join(ROOT, "oracle", "mexstub")])


def test_bindings_compile_and_link():
    assert os.path.exists(facade_exe()) and os.path.exists(shim_exe())


def test_bindings_fail_loudly_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([facade_exe()], capture_output=True, text=True)
    assert r.returncode == 100 and "CUDA" in r.stderr      # no CPU fallback behind the facade


@pytest.mark.gpu
def test_cpp_facade_reproduces_reference_demo(golden_dir, tmp_path):
    with open(os.path.join(golden_dir, "kat.json")) as f:
        k1 = json.load(f)["K1"]["bytes"]
    r = subprocess.run([facade_exe()], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == k1                            # ab b1 2f 09, SimpleCABAC.cpp's str.bin
    fn = str(tmp_path / "str.bin")
    r = subprocess.run([facade_exe(), fn], capture_output=True, text=True)
    assert r.returncode == 0 and open(fn, "rb").read().hex() == k1


@pytest.mark.gpu
def test_mex_shim_protocol_and_batch_commands(tmp_path):
    r = subprocess.run([shim_exe(), str(tmp_path / "k2.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
