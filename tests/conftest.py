"""Shared fixtures.  Tests marked `gpu` need a B200; everything else runs on CPU.

Nothing here (or in any `-m gpu` test) reads /root/reference at run time: the
reference's vectors are the xz-compressed copies under tests/golden/vectors and
the compiled reference, when available, is the prebuilt oracle/_ref/.
"""
import hashlib
import json
import lzma
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def sha1(b) -> str:
    return hashlib.sha1(bytes(b)).hexdigest()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def vectors(golden):
    """name -> bytes of the reference's test vectors (decompressed, SHA-1 checked)."""
    out = {}
    vdir = os.path.join(GOLDEN_DIR, "vectors")
    for fn in sorted(os.listdir(vdir)):
        name = fn[:-3]
        raw = lzma.decompress(open(os.path.join(vdir, fn), "rb").read())
        assert sha1(raw) == golden["vector_sha1"][name], name
        out[name] = raw
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled to oracle/_ref (skips when absent)."""
    from oracle import binding
    if not binding.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return binding.reference_lib()


@pytest.fixture(scope="session")
def lib():
    """The product C-ABI library.  Missing library is an error, never a skip."""
    import bjxa_b200
    return bjxa_b200.load()
