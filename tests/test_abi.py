"""The C-ABI library loads on a machine without a GPU and exports every symbol
that include/*.h declares, under the reference's ELF version nodes
(/root/reference/src/libbjxa.map:16-47).  No compute calls here."""
import os
import re
import subprocess

import bjxa_b200
from bjxa_b200.api import BATCH_SYMBOLS
from bjxa_b200.capi import SYMBOLS
from conftest import ROOT


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(bjxa_[a-z0-9_]+)\s*\(", text))


def test_headers_and_bindings_agree():
    assert declared("bjxa.h") == set(SYMBOLS)
    assert declared("bjxa_batch.h") == set(BATCH_SYMBOLS)
    assert len(SYMBOLS) == 19


def test_library_exports_every_declared_symbol(lib):
    for name in list(SYMBOLS) + list(BATCH_SYMBOLS):
        assert getattr(lib.dll, name) is not None, name


def test_symbol_versions_match_reference_map():
    out = subprocess.run(["readelf", "--dyn-syms", "-W", bjxa_b200.LIB_PATH],
                         capture_output=True, text=True, check=True).stdout
    ver = dict(re.findall(r"\b(bjxa_\w+)@@(\w[\w.]*)", out))
    node_01 = {"bjxa_decode", "bjxa_decode_format", "bjxa_decoder", "bjxa_dump_pcm",
               "bjxa_dump_riff_header", "bjxa_fread_header", "bjxa_free_decoder",
               "bjxa_fwrite_pcm", "bjxa_fwrite_riff_header", "bjxa_parse_header"}
    node_05 = {"bjxa_dump_header", "bjxa_encode", "bjxa_encode_format",
               "bjxa_encode_init", "bjxa_encoder", "bjxa_fread_riff_header",
               "bjxa_free_encoder", "bjxa_fwrite_header", "bjxa_parse_riff_header"}
    assert {k for k, v in ver.items() if v == "LIBBJXA_0.1"} == node_01
    assert {k for k, v in ver.items() if v == "LIBBJXA_0.5"} == node_05
    assert {k for k, v in ver.items() if v == "LIBBJXA_B200_1.0"} == set(BATCH_SYMBOLS)
    # nothing else leaks out of the library
    defined = re.findall(r"(?:FUNC|OBJECT)\s+GLOBAL\s+DEFAULT\s+\d+\s+(\w+)", out)
    assert defined and all(o.startswith("bjxa_") for o in defined), defined


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under bjxa_b200/ may import,
    link or execute it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bjxa_b200")):
        for fn in files:
            if fn.endswith((".py", ".c", ".cu", ".h", ".cc", ".map")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text, os.path.join(dirpath, fn)
    deps = subprocess.run(["ldd", bjxa_b200.LIB_PATH], capture_output=True,
                          text=True).stdout
    assert "oracle" not in deps and "bjxa_ref" not in deps


def test_installed_names_match_the_reference():
    """What the reference installs (/root/reference/Makefile.am:28,43): a library
    whose soname is libbjxa.so.0, libbjxa.so for -lbjxa, and bjxa.pc.  A program
    already linked against the reference finds the backend through the library
    path: no relink."""
    libdir = os.path.join(ROOT, "bjxa_b200", "lib")        # (not LIB_PATH: that may be another build)
    so0 = os.path.join(libdir, "libbjxa.so.0")
    dyn = subprocess.run(["readelf", "-d", so0], capture_output=True, text=True, check=True).stdout
    assert "Library soname: [libbjxa.so.0]" in dyn
    syms = lambda p: set(re.findall(r"\b(bjxa_\w+@@[\w.]+)", subprocess.run(      # noqa: E731
        ["readelf", "--dyn-syms", "-W", p], capture_output=True, text=True, check=True).stdout))
    assert syms(so0) == syms(os.path.join(libdir, "libbjxa_b200.so")) == syms(os.path.join(libdir, "libbjxa.so"))
    pc = open(os.path.join(libdir, "pkgconfig", "bjxa.pc")).read()
    assert "Name: bjxa" in pc and "-lbjxa" in pc and "@" not in pc
    # the reference CLI, dynamically linked against the REFERENCE's library
    # (oracle/_ref/refso, built by `make dropin`), resolves libbjxa.so.0 to ours
    # when the library path says so
    exe = os.path.join(ROOT, "oracle", "_ref", "bjxa_ref_dyn")
    if os.path.exists(exe):
        out = subprocess.run(["ldd", exe], capture_output=True, text=True,
                             env=dict(os.environ, LD_LIBRARY_PATH=libdir)).stdout
        assert os.path.join(libdir, "libbjxa.so.0") in out, out
