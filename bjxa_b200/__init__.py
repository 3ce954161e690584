"""bjxa_b200 -- B200-native batched BandJAM XA ADPCM block transform.

The product is the C-ABI shared library bjxa_b200/lib/libbjxa_b200.so (CUDA
kernels for sm_100a + the host layer in C behind the reference's bjxa.h API).
This package only locates and binds it; there is no Python or CPU fallback --
`load()` raises if the library has not been built.
"""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BJXA_B200_LIB: another build of the same library (`make sanitize` points the CPU
# test suite at one whose host C is compiled with -fsanitize=address,undefined)
LIB_PATH = os.environ.get("BJXA_B200_LIB") or os.path.join(_HERE, "lib", "libbjxa_b200.so")

_lib = None


def load():
    """Bind the product library (built by __graft_entry__.build())."""
    global _lib
    if _lib is None:
        from .api import Bjxa
        _lib = Bjxa(LIB_PATH)
    return _lib
