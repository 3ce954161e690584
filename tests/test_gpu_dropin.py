"""The drop-in acceptance test: the reference's OWN programs -- its CLI
(src/bjxa.c, bjxa_decode.c, bjxa_encode.c) and its API test
(test/test_libbjxa_api.c) -- compiled unmodified from /root/reference and
linked against the product library (`make dropin`, binaries under oracle/_ref,
prebuilt because /root/reference does not exist on the GPU box)."""
import os
import subprocess

import pytest

from conftest import ROOT, sha1

pytestmark = pytest.mark.gpu

REFDIR = os.path.join(ROOT, "oracle", "_ref")


def need(name):
    path = os.path.join(REFDIR, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not prebuilt (needs /root/reference at build time)")
    return path


XA_VECTORS = ["square-stereo-8.xa", "square-mono-8.xa", "square-stereo-6.xa",
              "square-mono-6.xa", "square-stereo-4.xa", "square-mono-4.xa"]


@pytest.mark.parametrize("binary", ["bjxa_dropin_single_pass", "bjxa_dropin"])
@pytest.mark.parametrize("name", XA_VECTORS)
def test_reference_cli_decode_hashes(vectors, golden, binary, name):
    """/root/reference/test/test_decode.sh:24-78 with `bjxa` = reference CLI + our
    lib: all six vectors, single pass and the CLI's default block-at-a-time mode
    (src/bjxa_decode.c:102-161: one bjxa_decode call per block, 10 000 to 20 000
    calls a vector, each one launch on the small-call path)."""
    exe = need(binary)
    r = subprocess.run([exe, "decode"], input=vectors[name], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert sha1(r.stdout) == golden["reference_tests"][name]["wav_sha1"]


def _hdr(magic=b"KWD1", data_len=682176, samples=661500, rate=44100, bits=8, ch=1):
    import struct
    return struct.pack("<4sIIHBBIhhhhI", magic, data_len, samples, rate, bits, ch, 0, 0, 0, 0, 0, 0)


# /root/reference/test/test_decode_error.sh:20-219, the message each case must give
CLI_HEADER_ERRORS = {
    "empty file": b"",
    "wrong magic number": _hdr(magic=b"KWD2"),
    "EIO (nDataLen 0)": _hdr(data_len=0),
    "ENOSAMPLES": _hdr(samples=0),
    "ETOOMANYSAMPLES": _hdr(data_len=33, samples=33),
    "ENOTENOUGHSAMPLES": _hdr(data_len=132, samples=32),
    "ENORATE": _hdr(rate=0),
    "data length not a multiple of the block size": _hdr(data_len=67, samples=64),
    "invalid number of bits": _hdr(bits=12),
    "invalid number of channels": _hdr(ch=5),
    "truncated header": _hdr()[:20],
}


@pytest.mark.parametrize("binary", ["bjxa_dropin_single_pass", "bjxa_dropin"])
@pytest.mark.parametrize("case", sorted(CLI_HEADER_ERRORS))
def test_reference_cli_header_errors(binary, case):
    exe = need(binary)
    r = subprocess.run([exe, "decode"], input=CLI_HEADER_ERRORS[case], capture_output=True, timeout=120)
    assert r.returncode != 0 and b"bjxa_fread_header" in r.stderr, (case, r.stderr)


@pytest.mark.parametrize("binary", ["bjxa_dropin_single_pass", "bjxa_dropin"])
def test_reference_cli_bad_block_profiles(binary):
    """/root/reference/test/test_decode_error.sh:221-282, byte for byte: an invalid
    mono block profile, an invalid right-channel block profile."""
    exe = need(binary)
    mono = _hdr(data_len=25, samples=32, bits=6, ch=1) + b"\xff" + bytes(24)
    right = _hdr(data_len=50, samples=32, bits=6, ch=2) + b"\x00" + bytes(24) + b"\xff" + bytes(24)
    for xa in (mono, right):
        r = subprocess.run([exe, "decode"], input=xa, capture_output=True, timeout=120)
        assert r.returncode != 0 and b"bjxa_decode" in r.stderr, r.stderr


def test_reference_cli_argument_errors(tmp_path):
    """/root/reference/test/test_bjxa.sh:57-88 (the CLI's own argument handling,
    unchanged by the library underneath)."""
    exe = need("bjxa_dropin")
    for args, msg in (
            (["decode", "src.xa", "dst.wav", "jnk.arg"], b"Too many arguments"),
            (["encode", "--bits", "4", "src.xa", "dst.wav", "jnk.arg"], b"Too many arguments"),
            (["decode", str(tmp_path / "nonexistent.xa")], b"Error:"),
            (["encode", "--bits"], b"Missing number of bits per sample"),
            (["encode", "--bits", "5"], b"Invalid number of bits per sample")):
        r = subprocess.run([exe] + args, capture_output=True, timeout=120, stdin=subprocess.DEVNULL)
        assert r.returncode != 0 and msg in r.stderr + r.stdout, (args, r.stderr, r.stdout)


def test_reference_cli_encode_block_at_a_time(vectors, golden):
    """The CLI's default encode loop, one bjxa_encode per block (src/bjxa_encode.c)."""
    exe = need("bjxa_dropin")
    for bits, wav in ((4, "square-stereo.wav"), (8, "square-mono.wav")):
        r = subprocess.run([exe, "encode", "--bits", str(bits)], input=vectors[wav],
                           capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert sha1(r.stdout) == golden["derived"]["encode_sha1"][f"{wav}:{bits}"]


def test_reference_cli_encode(vectors, golden):
    exe = need("bjxa_dropin_single_pass")
    for bits in (4, 6, 8):
        r = subprocess.run([exe, "encode", "--bits", str(bits)], input=vectors["square-mono.wav"],
                           capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert sha1(r.stdout) == golden["derived"]["encode_sha1"][f"square-mono.wav:{bits}"]


def test_reference_cli_bad_profile(vectors):
    """/root/reference/test/test_decode_error.sh:221-282"""
    exe = need("bjxa_dropin_single_pass")
    xa = bytearray(vectors["square-mono-8.xa"])
    xa[32 + 33 * 5] = 0xFF
    r = subprocess.run([exe, "decode"], input=bytes(xa), capture_output=True, timeout=600)
    assert r.returncode != 0 and b"bjxa_decode" in r.stderr


def test_reference_api_test_program(vectors, tmp_path):
    """/root/reference/test/test_libbjxa_api.c, unmodified, against our lib."""
    exe = need("test_api_dropin")
    (tmp_path / "test").mkdir()
    (tmp_path / "test" / "square-mono-4.xa").write_bytes(vectors["square-mono-4.xa"])
    r = subprocess.run([exe], env=dict(os.environ, SRCDIR=str(tmp_path)),
                       capture_output=True, timeout=600, stdin=subprocess.DEVNULL)
    assert r.returncode == 0, (r.returncode, r.stderr[-500:])


@pytest.mark.parametrize("name", ["square-stereo-4.xa", "square-mono-6.xa"])
def test_no_relink(vectors, golden, name):
    """A program ALREADY linked against the reference's libbjxa.so.0 (its CLI, built
    against oracle/_ref/refso): with the reference's library on the path it runs on
    the CPU, with ours on the path it runs on the B200 -- same bytes, no relink."""
    exe = need("bjxa_ref_dyn")
    libdirs = {"reference": os.path.join(REFDIR, "refso"),
               "b200": os.path.join(ROOT, "bjxa_b200", "lib")}
    for who, libdir in libdirs.items():
        env = dict(os.environ, LD_LIBRARY_PATH=libdir)
        ldd = subprocess.run(["ldd", exe], capture_output=True, text=True, env=env).stdout
        assert os.path.join(libdir, "libbjxa.so.0") in ldd, (who, ldd)
        r = subprocess.run([exe, "decode"], input=vectors[name], capture_output=True,
                           timeout=600, env=env)
        assert r.returncode == 0, (who, r.stderr)
        assert sha1(r.stdout) == golden["reference_tests"][name]["wav_sha1"], who
