"""gridtools_b200 -- B200 (sm_100a) stencil execution backend + halo exchange behind the GridTools interfaces.

The product is the C-ABI library `libgtb200.so` (include/gtb200.h, sources in gridtools_b200/csrc) and the
GridTools backend tag in include/gtb200/stencil/b200.hpp.  This package is the Python host used by the tests and
the benchmark: storage builder, stencil entry points and the gcl-style halo exchange.  Nothing here computes on
the CPU and nothing falls back to PyTorch ops.
"""
__version__ = "0.1.0"
