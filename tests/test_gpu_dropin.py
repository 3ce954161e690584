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


@pytest.mark.parametrize("binary", ["bjxa_dropin_single_pass", "bjxa_dropin"])
@pytest.mark.parametrize("name", ["square-stereo-4.xa", "square-mono-8.xa", "square-stereo-6.xa"])
def test_reference_cli_decode_hashes(vectors, golden, binary, name):
    """/root/reference/test/test_decode.sh with `bjxa` = reference CLI + our lib."""
    exe = need(binary)
    if binary == "bjxa_dropin" and name != "square-stereo-4.xa":
        pytest.skip("block-at-a-time mode is one GPU round trip per block; one vector is enough")
    r = subprocess.run([exe, "decode"], input=vectors[name], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert sha1(r.stdout) == golden["reference_tests"][name]["wav_sha1"]


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
