"""INTEGRATION.md section B, checked: tests/native/reference_bridge.cpp -- the binding a Movi maintainer adds to call
include/movi_gpu.h from inside the reference -- compiles against the reference's own headers for every index type.  Runs where
the reference is mounted (the build container); skipped elsewhere."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MOVI_REFERENCE", "/root/reference")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include")) or not CXX, reason="reference headers not mounted")
@pytest.mark.parametrize("mode", [6, 8, 7, 3, 2, 5])
def test_reference_side_binding_compiles_against_the_reference_headers(mode):
    r = subprocess.run([CXX, "-std=c++20", "-fsyntax-only", "-w", "-fopenmp", "-DCOLOR_MODE=0", "-DMODE=%d" % mode,
                        "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(REF, "include"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "native", "reference_bridge.cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
