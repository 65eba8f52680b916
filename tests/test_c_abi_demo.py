"""The C ABI from plain C (examples/c_abi_demo.c): include/ccvsq.h is valid C99 with no torch / C++ types, and a C
host program that owns its device buffers gets the reference's results through ONE call per forward / decode."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "ccvsq.h")
DEMO = os.path.join(ROOT, "examples", "c_abi_demo.c")
LIB_DIR = os.path.join(ROOT, "ccvs_b200", "lib")


def _cuda_home():
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "include", "cuda_runtime_api.h")):
            return c
    return None


def test_header_is_plain_c99():
    assert shutil.which("gcc"), "gcc is part of the image"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                   check=True, capture_output=True)


def test_c_demo_compiles_against_the_header():
    cuda = _cuda_home()
    if cuda is None:
        pytest.skip("no CUDA toolkit headers")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", f"-I{cuda}/include", DEMO],
                   check=True, capture_output=True)


@pytest.mark.gpu
def test_c_demo_runs_and_matches_the_host_search(tmp_path):
    cuda = _cuda_home()
    assert cuda is not None
    exe = str(tmp_path / "c_abi_demo")
    subprocess.run(["gcc", "-std=c99", "-O2", f"-I{cuda}/include", DEMO, "-o", exe, f"-L{LIB_DIR}", "-lccvsq",
                    f"-L{cuda}/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{LIB_DIR}", f"-Wl,-rpath,{cuda}/lib64"],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK"), r.stdout
