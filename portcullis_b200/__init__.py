"""portcullis_b200 — B200-native implementation of the Portcullis `junc` stage.

The product is the CUDA/C++ shared library ``libportcullis_junc.so`` (sources in ``csrc/``, C ABI in
``include/``).  This package is the thin Python mirror of the reference's JunctionBuilder interface
used by the tests and the benchmark harness.
"""
from . import _lib  # noqa: F401
from .junction_builder import JunctionBuilder, JuncGpu, PrepDir  # noqa: F401
