"""Runs the C++ tests of the host layer (gnuradio4_b200/host: gr::Block / Graph / scheduler::Simple mirror).
qa_plumbing is host-only (BASELINE config #1 + settings / seam / connect behaviour); qa_device_graph needs the B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "cpp")


def ensure_built():
    lib = os.path.join(ROOT, "gnuradio4_b200", "libgr4b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "gnuradio4_b200", "csrc"), "-j8"], check=True, capture_output=True)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True, capture_output=True)
    result = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], capture_output=True, text=True)
    assert result.returncode == 0, result.stdout + result.stderr


def test_qa_plumbing_host_only():
    ensure_built()
    result = subprocess.run([os.path.join(BIN, "qa_plumbing")], capture_output=True, text=True, timeout=300)
    print(result.stdout)
    assert result.returncode == 0, result.stdout + result.stderr
    assert "config1_plumbing_msamples_per_s" in result.stdout


@pytest.mark.gpu
def test_qa_device_graph():
    binary = os.path.join(BIN, "qa_device_graph")
    if not os.path.exists(binary):
        ensure_built()
    result = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    print(result.stdout)
    assert result.returncode == 0, result.stdout + result.stderr
    assert "FAIL" not in result.stdout
