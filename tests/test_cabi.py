"""The C-ABI library loads without a GPU and exports every symbol include/gr4b200.h declares; the host-side design
functions (no GPU needed) match the oracle bit for bit; the Python host layer mirrors the reference's error behaviour."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gr4():
    import gnuradio4_b200 as g

    if not os.path.exists(g._lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    g.load()
    return g


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gr4b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gr4b200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(gr4):
    import ctypes

    lib = ctypes.CDLL(gr4._lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 55
    for name in names:
        assert hasattr(lib, name), f"{name} declared in gr4b200.h but not exported"
        assert name in gr4._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(gr4._lib.SIGNATURES) == names
    assert gr4.load().gr4b200_abi_version() == 1
    assert gr4.load().gr4b200_device_count() >= 0  # 0 on the CPU container, no crash


def test_design_functions_match_oracle(gr4, oracle):
    from tests import _oracle

    for name in _oracle.WINDOWS:
        for n in (2, 8, 127, 4096):
            assert np.array_equal(gr4.window(name, n).view(np.uint32), oracle.window(name, n).view(np.uint32)), (name, n)
    for nt, fc, win in ((127, 0.1, "Hamming"), (127, 0.05, "Hamming"), (3072, 1 / 512, "Kaiser"), (64, 0.3, "Blackman")):
        assert np.array_equal(gr4.fir_generate(nt, win, fc).view(np.uint32), oracle.fir_generate(nt, win, fc).view(np.uint32))
    for ftype in ("LOWPASS", "HIGHPASS", "BANDPASS", "BANDSTOP"):
        for win in ("Kaiser", "Hamming", "Hann"):
            a = gr4.fir_design(ftype, 4, 1.0, 10.0, 1000.0, window_type=win)
            assert np.array_equal(a.view(np.uint32), oracle.fir_design(ftype, 4, 1.0, 10.0, 1000.0, window=win).view(np.uint32)), (ftype, win)
    with pytest.raises(gr4.Gr4b200Error):
        gr4.window("Kaiser", 8, beta=-1.0)


def test_compute_domain_grammar(gr4):
    """core/test/qa_ComputeDomain.cpp:186-189 style: kind[:backend[:index]], host aliases."""
    from gnuradio4_b200.blocks import parse_compute_domain

    assert parse_compute_domain("gpu:cuda:3") == ("gpu", "cuda", 3)
    assert parse_compute_domain("gpu") == ("gpu", "sycl", -1)
    assert parse_compute_domain("gpu:cuda") == ("gpu", "cuda", -1)
    assert parse_compute_domain("gpu:cuda:x") == ("gpu", "cuda", -1)
    for alias in ("", "host", "default_cpu", "default_io", "weird"):
        assert parse_compute_domain(alias) == ("host", "none", -1)
    with pytest.raises(gr4.Gr4b200Error):
        gr4.MultiplyConst(value=2, compute_domain="host")  # no host path in this package: fail loudly


def test_product_does_not_touch_the_oracle():
    """The product path must not import, link or load anything under oracle/ (checked statically)."""
    pkg = os.path.join(ROOT, "gnuradio4_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "libgr4ref" not in text and "oracle/" not in text and "_oracle" not in text, os.path.join(dirpath, f)


def test_graph_rejects_non_chains(gr4):
    from gnuradio4_b200.flowgraph import Graph

    class Dummy:
        input_chunk_size = output_chunk_size = 1

    g = Graph()
    a, b, c = (g.emplaceBlock(Dummy) for _ in range(3))
    assert g.connect(a, b)
    with pytest.raises(gr4.Gr4b200Error):
        g.connect(a, c)  # second edge from the same output port
    with pytest.raises(gr4.Gr4b200Error):
        g.chain()  # c is unconnected
    assert g.connect(b, c) and g.chain() == [a, b, c]


def test_grc_documents_parse_like_the_reference_importer(gr4):
    """Graph_yaml_importer.hpp:88-380: blocks with id / parameters.name, connections of >= 4 elements, type-tagged values."""
    text = """
blocks:
  - id: gr::filter::fir_filter<complex64>
    parameters:
      name: lowpass
      b: [0.25, 0.5, 0.25]
  - id: gr::blocks::fft::FFT<complex64>
    parameters:
      name: spectrum
      fftSize: !!uint32 4096
      sample_rate: !!float32 1000
connections:
  - [lowpass, 0, spectrum, 0, 8192]
"""
    blocks, connections = gr4.parse_grc(text)
    assert blocks == [("gr::filter::fir_filter", "complex64", "lowpass", {"b": [0.25, 0.5, 0.25]}), ("gr::blocks::fft::FFT", "complex64", "spectrum", {"fftSize": 4096, "sample_rate": 1000.0})]
    assert connections == [("lowpass", 0, "spectrum", 0, 8192)]
    with pytest.raises(gr4.Gr4b200Error):
        gr4.parse_grc("blocks:\n  - id: x\n    parameters: {}\n")  # no name
    with pytest.raises(gr4.Gr4b200Error):
        gr4.parse_grc("blocks: []\nconnections:\n  - [a, 0, b]\n")  # three elements
