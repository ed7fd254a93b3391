"""The C++ facade (include/mci/*.hpp over the C-ABI): compile the reference-style test program and example with g++ and run
them — the API-only part on CPU, the full statistical / replay checks on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mcintegratorplusplus_b200")


def _compile(src, out):
    cc = ["gcc", "-std=c99"] if src.endswith(".c") else ["g++", "-std=c++14"]
    cmd = cc + ["-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), src, "-L" + PKG, "-lmcig", "-lm",
                "-Wl,-rpath," + PKG, "-o", out]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def test_facade_compiles_and_api_behaviour(mcig, tmp_path):
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"), str(tmp_path / "test_facade"))
    out = subprocess.run([exe, "api"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "api ok" in out.stdout
    _compile(os.path.join(ROOT, "examples", "ex_basic.cpp"), str(tmp_path / "ex_basic"))


@pytest.mark.gpu
def test_facade_full_on_gpu(mcig, tmp_path):
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"), str(tmp_path / "test_facade"))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu ok" in out.stdout


@pytest.mark.gpu
def test_ex_basic_on_gpu(mcig, tmp_path):
    exe = _compile(os.path.join(ROOT, "examples", "ex_basic.cpp"), str(tmp_path / "ex_basic"))
    out = subprocess.run([exe, "512"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")


def test_plain_c_binding_api(mcig, tmp_path):
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_capi.c"), str(tmp_path / "test_capi"))
    out = subprocess.run([exe, "api"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "capi api ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_plain_c_binding_on_gpu(mcig, tmp_path):
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_capi.c"), str(tmp_path / "test_capi"))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "capi gpu ok" in out.stdout, out.stdout + out.stderr
