"""The drop-in claim, from C++: examples/octant_cycle.cpp uses nothing but the
TMROctForest class API.  Compiled against the REFERENCE's headers + the oracle
library it prints the lines committed under tests/golden/octant_cycle_*.txt;
compiled against this repository's headers it must print the same lines --
with the test-only emulation here, with the CUDA library under -m gpu.
(Named zz so that it runs after the parity tests.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "octant_cycle.cpp")
HOST = os.path.join(ROOT, "tmr_b200", "csrc", "host")
REF_SRC = "/root/reference/src"

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")


def _build(tmp_path, name, includes, libdir, lib):
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++14", "-O1"] + ["-I" + i for i in includes] + [
        "-I" + os.path.join(ROOT, "include"), SRC, "-L" + libdir, "-l" + lib,
        "-Wl,-rpath," + libdir, "-pthread", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _run(exe, order):
    return subprocess.run([exe, str(order)], check=True, capture_output=True, text=True,
                          timeout=300).stdout


def _golden(order):
    return open(os.path.join(ROOT, "tests", "golden", "octant_cycle_order%d.txt" % order)).read()


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference headers not present")
def test_same_source_against_reference_headers(tmp_path, ref_lib):
    """the committed golden lines ARE what the unmodified reference prints"""
    exe = _build(tmp_path, "cycle_ref", [os.path.join(ROOT, "oracle", "shim"), REF_SRC],
                 os.path.join(ROOT, "oracle", "_ref"), "tmr_ref")
    for order in (2, 3):
        assert _run(exe, order) == _golden(order)


def test_same_source_against_dropin_headers(tmp_path, emu_lib):
    exe = _build(tmp_path, "cycle_emu", [HOST, os.path.join(HOST, "shim")],
                 os.path.join(ROOT, "tests", "emu", "_build"), "tmr_emu")
    for order in (2, 3):
        assert _run(exe, order) == _golden(order)


@pytest.mark.gpu
def test_same_source_on_the_cuda_library(tmp_path, gpu_lib):
    exe = _build(tmp_path, "cycle_gpu", [HOST, os.path.join(HOST, "shim")],
                 os.path.join(ROOT, "tmr_b200", "lib"), "tmr_b200")
    for order in (2, 3):
        assert _run(exe, order) == _golden(order)
