"""Whole files through bjxa_corpus_run (include/bjxa_batch.h): the container step
on either side of the block transform -- header parse / validation, header
emit, file assembly on the device.  The authorities are the reference's own
whole-file goldens (test/test_decode.sh:24-78), the derived encode goldens and
the oracle's file-level functions."""
import errno

import numpy as np
import pytest

from bjxa_b200 import synth
from bjxa_b200.api import CORPUS_WAV_TO_XA, CORPUS_XA_TO_WAV
from conftest import sha1

pytestmark = pytest.mark.gpu

XA_VECTORS = ["square-stereo-8.xa", "square-mono-8.xa", "square-stereo-6.xa",
              "square-mono-6.xa", "square-stereo-4.xa", "square-mono-4.xa"]


def test_reference_vectors_as_one_corpus(lib, vectors, golden):
    """The six shipped .xa files in one call: every WAV hashes to the value the
    reference's test suite expects."""
    table, wavs = lib.corpus(CORPUS_XA_TO_WAV, [vectors[n] for n in XA_VECTORS])
    assert (table["error"] == 0).all()
    for name, t, wav in zip(XA_VECTORS, table, wavs):
        assert sha1(wav) == golden["reference_tests"][name]["wav_sha1"], name
        assert t["out_off"] % 16 == 4 and t["channels"] == (2 if "stereo" in name else 1)
        assert t["bits"] == int(name[-4]) and t["rate"] == 44100


def test_source_wavs_as_one_corpus(lib, vectors, golden):
    for bits in (4, 6, 8):
        names = ["square-mono.wav", "square-stereo.wav"]
        table, xas = lib.corpus(CORPUS_WAV_TO_XA, [vectors[n] for n in names], bits=bits)
        assert (table["error"] == 0).all()
        for name, xa in zip(names, xas):
            assert sha1(xa) == golden["derived"]["encode_sha1"][f"{name}:{bits}"]
        # XA files are packed back to back
        assert table[1]["out_off"] == table[0]["out_off"] + table[0]["out_len"]


def _xa_file(seed, key, bits, ch, samples, mix="P1", prev=((0, 0), (0, 0))):
    return synth.make_xa(seed, key, bits, ch, samples, mix=mix, prev=prev)


def test_mixed_corpus_with_rejects(lib, oracle):
    """Good files of every shape between files the reference rejects: each good
    one equals the oracle's WAV, each bad one carries the reference's errno, and
    nothing is written outside the span of the produced files (the padding
    between two WAV files travels with the chunk's single copy: unspecified)."""
    files, kinds = [], []
    k = 0
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for samples in (1, 31, 32, 33, 5000, 32 * 700 + 9):
                k += 1
                files.append(_xa_file(77, k, bits, ch, samples, mix=("P0", "P1", "P2", "P3")[k % 4],
                                      prev=((k, -k), (3 * k, 7))))
                kinds.append("ok")
            good = _xa_file(78, k, bits, ch, 3200)
            files.append(b"KWD2" + good[4:])                      # bad magic
            kinds.append("header")
            files.append(good[:len(good) - 5])                    # payload cut short
            kinds.append("short")
            files.append(good[:20])                               # not even a header
            kinds.append("short")
            bad = bytearray(_xa_file(79, k, bits, ch, 32 * 50, mix="P0"))
            bad[32 + (4 * bits + 1) * ch * 20] = 0x7f             # filter 7 in block 20
            files.append(bytes(bad))
            kinds.append("profile")
    table, wavs = lib.corpus(CORPUS_XA_TO_WAV, files)
    out = lib.last_corpus_out
    ends = []
    for f, kind, t, wav in zip(files, kinds, table, wavs):
        want = oracle.xa_to_wav(f)
        if kind == "ok":
            assert t["error"] == 0 and wav == want
        elif kind == "header":
            assert want is None and t["error"] == errno.EPROTO and t["out_len"] == 0
        elif kind == "short":
            assert t["error"] == errno.EIO and t["out_len"] == 0
        else:
            ch = int(t["channels"])
            assert t["error"] == errno.EPROTO and t["blocks"] == 20
            assert t["out_len"] == 44 + 20 * 64 * ch
            # the blocks in front of the bad one are the reference's
            good = bytearray(f)
            good[32 + (len(f) - 32) // 50 * 20] = 0
            assert wav[44:] == oracle.xa_to_wav(bytes(good))[44:44 + 20 * 64 * ch]
        if kind == "ok":
            ends.append(int(t["out_off"] + t["out_len"]))
        elif kind == "profile":     # the whole file's slot was laid out
            ends.append(int(t["out_off"]) + 44 + 64 * ch * 50)
    live = table[table["out_len"] > 0]
    assert (out[:int(live["out_off"].min())] == 0xCD).all()
    assert (out[max(ends):] == 0xCD).all()


def test_many_files_span_several_chunks(lib, oracle):
    """~70 MB of input: more than one chunk per pipeline slot."""
    protos = [_xa_file(91, i, (4, 6, 8)[i % 3], 1 + i % 2, 32 * (300 + 37 * i) + i,
                       mix=("P1", "P2")[i % 2]) for i in range(12)]
    files = [protos[i % 12] for i in range(4200)]
    assert sum(len(f) for f in files) > 64 << 20
    table, wavs = lib.corpus(CORPUS_XA_TO_WAV, files)
    assert (table["error"] == 0).all()
    want = [oracle.xa_to_wav(p) for p in protos]
    for i, wav in enumerate(wavs):
        assert wav == want[i % 12], i


def test_pipeline_slots_are_reused(lib, oracle, monkeypatch):
    """Tiny chunks: dozens of turns through the three pipeline slots, every plan
    reset and re-run many times."""
    monkeypatch.setenv("BJXA_B200_CORPUS_CHUNK", str(300 << 10))
    protos = [_xa_file(92, i, (4, 6, 8)[i % 3], 1 + i % 2, 32 * (900 + 211 * i) + i,
                       mix=("P1", "P2", "P3")[i % 3]) for i in range(9)]
    files = [protos[(i * 5) % 9] for i in range(400)]
    table, wavs = lib.corpus(CORPUS_XA_TO_WAV, files)
    assert (table["error"] == 0).all()
    want = [oracle.xa_to_wav(p) for p in protos]
    for i, wav in enumerate(wavs):
        assert wav == want[(i * 5) % 9], i


def test_wav_corpus_vs_oracle(lib, oracle):
    files = []
    for i, (bits, ch) in enumerate([(4, 1), (4, 2), (6, 1), (6, 2), (8, 1), (8, 2)] * 3):
        pcm = synth.make_pcm(55, i, ch, 1 + 997 * i)
        files.append((bits, synth.riff_header(pcm.size * 2, ch) + pcm.tobytes()))
    for bits in (4, 6, 8):
        sel = [f for b, f in files if b == bits]
        table, xas = lib.corpus(CORPUS_WAV_TO_XA, sel, bits=bits)
        assert (table["error"] == 0).all()
        for f, xa in zip(sel, xas):
            assert xa == oracle.wav_to_xa(f, bits)
        # the same files packed back to back: their PCM is no longer 16-byte
        # aligned in the arena and travels file by file
        table, xas = lib.corpus(CORPUS_WAV_TO_XA, sel, bits=bits, align=False)
        assert (table["error"] == 0).all() and (table["in_off"][1:] % 16 != 4).any()
        for f, xa in zip(sel, xas):
            assert xa == oracle.wav_to_xa(f, bits)
    # a WAV the reference rejects (8-bit PCM) and one cut short
    bad = bytearray(files[0][1])
    bad[34] = 8
    table, xas = lib.corpus(CORPUS_WAV_TO_XA, [bytes(bad), files[1][1][:100], files[2][1]], bits=4)
    assert table["error"].tolist()[:2] != [0, 0] and table[1]["error"] == errno.EIO
    assert table[2]["error"] == 0 and xas[2] == oracle.wav_to_xa(files[2][1], 4)
