"""gnuradio4_b200 -- B200-native hot path for GNU Radio 4 streaming DSP blocks (FIR, FFT, mixer, elementwise math).

The product is the C-ABI shared library `libgr4b200.so` (include/gr4b200.h, sources in gnuradio4_b200/csrc). This Python
layer is a thin host-side mirror of the reference's block interface for that path -- same block names, settings and
error behaviour as fair-acc/gnuradio4's blocks/{filter,fourier,math} -- used by tests/ and bench.py. PyTorch only supplies
device memory and streams. There is no CPU fallback.
"""
from . import _lib
from ._lib import Gr4b200Error, load
from .blocks import (FFT, AddConst, Add, BasicDecimatingFilter, ComplexToInterleaved, DDC, InterleavedToComplex, FirFft, Decimator, Divide, DivideConst, Multiply, MultiplyConst, PolyphaseChannelizer, PolyphaseResampler, Rotator, Subtract, SubtractConst, fir_filter, fir_design, fir_generate, window)
from .flowgraph import Graph, HostBuffer, Simple, load_grc, parse_grc, save_grc

__all__ = ["FFT", "AddConst", "Add", "BasicDecimatingFilter", "ComplexToInterleaved", "DDC", "InterleavedToComplex", "FirFft", "Decimator", "Divide", "DivideConst", "Multiply", "MultiplyConst", "PolyphaseChannelizer", "PolyphaseResampler", "Rotator", "Subtract", "SubtractConst", "fir_filter", "fir_design", "fir_generate", "window", "Graph", "HostBuffer", "Simple", "load_grc", "parse_grc", "save_grc", "Gr4b200Error", "load"]
