"""The oracle (oracle/bjxa_oracle.c) against every golden the reference's own
tests hold for the path, the derived goldens, and -- when oracle/_ref exists --
the compiled, unmodified reference on seeded streams.  CPU only."""
import numpy as np
import pytest

from bjxa_b200 import synth
from conftest import sha1

XA_VECTORS = ["square-stereo-8.xa", "square-mono-8.xa", "square-stereo-6.xa",
              "square-mono-6.xa", "square-stereo-4.xa", "square-mono-4.xa"]


def saturation_xa():
    """/root/reference/test/test_decode.sh:88-119"""
    return (synth.xa_header(66, 32, 44100, 8, 2)
            + b"\x20" + b"\x7f" * 32 + b"\x20" + b"\x80" * 32)


@pytest.mark.parametrize("name", XA_VECTORS)
def test_reference_decode_hashes(oracle, vectors, golden, name):
    """/root/reference/test/test_decode.sh:24-78"""
    g = golden["reference_tests"][name]
    assert sha1(vectors[name]) == g["input_sha1"]
    wav = oracle.xa_to_wav(vectors[name])
    assert sha1(wav) == g["wav_sha1"]
    assert sha1(wav[44:]) == golden["derived"]["pcm_sha1"][name]


def test_saturation_known_answer(oracle, golden):
    """/root/reference/test/test_decode.sh:80-122: L = 32512 then 31 x 32767,
    R = 32 x -32768."""
    wav = oracle.xa_to_wav(saturation_xa())
    assert sha1(wav) == golden["reference_tests"]["saturation"]["wav_sha1"]
    pcm = np.frombuffer(wav[44:], dtype="<i2").reshape(32, 2)
    assert pcm[0, 0] == 32512 and (pcm[1:, 0] == 32767).all()
    assert (pcm[:, 1] == -32768).all()


def test_semantics_probes(oracle):
    """SURVEY.md section 8c probes, measured on the compiled reference."""
    # header state (100,-50), filter 2, range 0, all-zero codes
    blk = np.zeros(33, dtype=np.uint8)
    blk[0] = 0x20
    done, bad, pcm, st = oracle.decode_blocks(8, 1, [[100, -50], [0, 0]], blk, 1, 64)
    assert done == 1 and not bad and list(pcm[:4]) == [220, 314, 385, 436]
    assert list(st[0]) == [pcm[31], pcm[30]]
    # prev0 = -1, filter 1: -240/256 truncates toward zero
    blk[0] = 0x10
    _, _, pcm, _ = oracle.decode_blocks(8, 1, [[-1, 0], [0, 0]], blk, 1, 64)
    assert pcm[0] == 0
    # range 15 keeps only the sign
    blk[0] = 0x0F
    blk[1:5] = [0x80, 0x7F, 0xFF, 0x01]
    _, _, pcm, _ = oracle.decode_blocks(8, 1, [[0, 0], [0, 0]], blk, 1, 64)
    assert list(pcm[:4]) == [-1, 0, -1, 0]


@pytest.mark.parametrize("chan", [0, 1])
def test_bad_profile_stops(oracle, chan):
    """/root/reference/test/test_decode_error.sh:221-282: profile 0xff in a
    mono block / in the right block of a stereo pair."""
    ch = chan + 1
    xa = np.frombuffer(synth.make_xa(1, 2, 8, ch, 96, "P2")[32:], dtype=np.uint8).copy()
    bad_at = 1                       # second effective block
    xa[(bad_at * ch + chan) * 33] = 0xFF
    done, bad, pcm, st = oracle.decode_blocks(8, ch, [[0, 0], [0, 0]], xa, 3, 96 * 2 * ch)
    assert bad and done == bad_at and pcm.size == bad_at * 32 * ch
    if chan == 1:   # the left channel HAS advanced over the failing pair
        _, _, p2, st2 = oracle.decode_blocks(8, 1, [[0, 0], [0, 0]],
                                             xa.reshape(-1, 33)[0::2].reshape(-1), 2, 128)
        assert list(st[0]) == [p2[63], p2[62]]


def test_derived_encode_hashes(oracle, vectors, golden):
    for name in ("square-mono.wav", "square-stereo.wav"):
        for bits in (4, 6, 8):
            xa = oracle.wav_to_xa(vectors[name], bits)
            assert sha1(xa) == golden["derived"]["encode_sha1"][f"{name}:{bits}"]
            nb = (len(xa) - 32) // (4 * bits + 1)
            prof = np.frombuffer(xa[32:], dtype=np.uint8).reshape(nb, -1)[:, 0]
            assert not prof.any()    # /root/reference/src/libbjxa.c:679


def test_differential_goldens_decode(oracle, golden):
    seed = golden["differential"]["seed"]
    for c in golden["differential"]["decode"]:
        xa = synth.make_xa(seed, c["key"], c["bits"], c["channels"], c["samples"],
                           c["mix"], c["prev"])
        assert sha1(xa) == c["input_sha1"], "synthetic generator drifted"
        assert sha1(oracle.xa_to_wav(xa)) == c["wav_sha1"], c


def test_differential_goldens_encode(oracle, golden):
    seed = golden["differential"]["seed"]
    for c in golden["differential"]["encode"]:
        pcm = synth.make_pcm(seed, c["key"], c["channels"], c["frames"])
        wav = synth.riff_header(pcm.size * 2, c["channels"]) + pcm.tobytes()
        assert sha1(wav) == c["input_sha1"], "synthetic generator drifted"
        assert sha1(oracle.wav_to_xa(wav, c["bits"])) == c["xa_sha1"], c


def test_roundtrip_encode_decode_top_bits(oracle):
    """decode(encode(x)) keeps exactly the top `bits` bits (profile 0)."""
    pcm = synth.make_pcm(3, 4, 2, 1000)
    for bits in (4, 6, 8):
        xa = oracle.encode_blocks(bits, 2, pcm)
        done, bad, out, _ = oracle.decode_blocks(bits, 2, [[0, 0], [0, 0]], xa,
                                                 (1000 + 31) // 32, pcm.size * 2)
        mask = np.int16(-(1 << (16 - bits)))
        assert not bad and np.array_equal(out, pcm & mask)


def test_oracle_vs_compiled_reference(oracle, ref):
    """Fresh seeds (not in golden.json), all mixes, through the real API."""
    key = 0
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for mix in synth.MIXES:
                key += 1
                xa = synth.make_xa(99, key, bits, ch, 3000 + key, mix,
                                   ((key, -key), (3 * key, 5)))
                assert oracle.xa_to_wav(xa) == ref.xa_to_wav(xa)
    pcm = synth.make_pcm(99, 7, 2, 4321)
    wav = synth.riff_header(pcm.size * 2, 2) + pcm.tobytes()
    for bits in (4, 6, 8):
        assert oracle.wav_to_xa(wav, bits) == ref.wav_to_xa(wav, bits)


def test_extremes_against_compiled_reference(oracle, ref):
    """The hand-made streams of batchgen.extremes (both rails, every filter, every
    sign of the truncating division) decode identically with the restatement and
    with the unmodified reference -- and do reach both rails."""
    import batchgen
    rails = set()
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for s in batchgen.extremes(bits, ch, blocks=12):
                pay = bytes(s["payload"])
                xa = synth.xa_header(len(pay), s["samples"], 44100, bits, ch, s["prev"]) + pay
                wav = oracle.xa_to_wav(xa)
                assert wav == ref.xa_to_wav(xa), (bits, ch, s["prev"])
                pcm = np.frombuffer(wav[44:], dtype="<i2")
                rails |= {int(pcm.min()), int(pcm.max())}
    assert {-32768, 32767} <= rails
