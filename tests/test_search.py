"""The searching encoder (BJXA_PLAN_ENCODE_SEARCH) -- an EXTENSION: the reference
encoder hard-codes profile 0 (src/libbjxa.c:679), so there is no reference
output to compare with.  What pins it:

  * the plain-C restatement in oracle/bjxa_oracle.c, which decodes every block
    it emits with the (pinned) decoder restatement and refuses to return if the
    decoder disagrees with the search's own reconstruction;
  * properties checked here against the UNMODIFIED reference decoder: the output
    decodes without error, and its squared error never exceeds that of the
    reference encoder's output -- per block, not just in total;
  * the product's candidate arithmetic (xa_core.h), stepped on the CPU by the
    emulator, equal to the oracle bit for bit.
"""
import numpy as np
import pytest

import batchgen
from bjxa_b200 import synth
from oracle import binding

ZERO = [[0, 0], [0, 0]]


@pytest.fixture(scope="module")
def oracle():
    return binding.Oracle()


def _block_errors(x, y, channels):
    """squared error per block-channel: (blocks, channels)"""
    frames = x.size // channels
    blocks = (frames + 31) // 32
    e = np.zeros((blocks * 32, channels), dtype=np.int64)
    d = x.reshape(-1, channels).astype(np.int64) - y.reshape(-1, channels).astype(np.int64)
    e[:frames] = d * d
    return e.reshape(blocks, 32, channels).sum(axis=1)


def _ref_decode(bits, channels, xa, frames, rate=44100):
    """through the unmodified reference library when it was built, else the oracle"""
    blocks = (frames + 31) // 32
    if binding.have_ref():
        lib = binding.reference_lib()
        dec = lib.decoder()
        hdr = synth.xa_header(xa.size, frames, rate, bits, channels)
        lib.parse_header(dec, hdr)
        out = np.zeros(blocks * 64 * channels, dtype=np.uint8)
        got = lib.decode(dec, out, out.size, xa, xa.size)
        lib.free_decoder(dec)
        assert got == blocks
        return out.view(np.int16)[:frames * channels]
    done, bad, pcm, _ = binding.Oracle().decode_blocks(bits, channels, ZERO, xa, blocks,
                                                       frames * channels * 2)
    assert done == blocks and not bad
    return pcm


@pytest.mark.parametrize("bits", [4, 6, 8])
@pytest.mark.parametrize("channels", [1, 2])
def test_oracle_search_decodes_with_the_reference_and_never_loses(oracle, bits, channels):
    frames = 32 * 90 + 7
    pcm = synth.make_pcm(11, bits * 10 + channels, channels, frames)
    xa, state = oracle.encode_search_blocks(bits, channels, ZERO, pcm)
    back = _ref_decode(bits, channels, xa, frames)
    plain = _ref_decode(bits, channels, oracle.encode_blocks(bits, channels, pcm), frames)
    e_search = _block_errors(pcm, back, channels)
    e_plain = _block_errors(pcm, plain, channels)
    assert (e_search <= e_plain).all()
    assert e_search.sum() * 4 < e_plain.sum()       # and it is worth having: > 6 dB here
    profiles = xa[::4 * bits + 1]
    assert (profiles >> 4).max() <= 4 and ((profiles & 15) <= 16 - bits).all()
    assert len(set((profiles >> 4).tolist())) >= 3   # it does use the filters
    # the state it reports is the decoder's: last two samples of each channel
    last = back.reshape(-1, channels)
    pad = (-frames) % 32                             # the zero padding was encoded as well
    if pad == 0:
        for c in range(channels):
            assert state[c] == [int(last[-1, c]), int(last[-2, c])]


@pytest.mark.parametrize("bits", [4, 6, 8])
def test_oracle_search_known_answers(oracle, bits):
    # silence from a silent state: every candidate is exact, the lowest profile byte wins
    xa, state = oracle.encode_search_blocks(bits, 1, ZERO, np.zeros(64, dtype=np.int16))
    assert not xa.any() and state[0] == [0, 0]
    # a constant that only the finest step can hit exactly: filter 0, largest range
    v = 3
    xa, _ = oracle.encode_search_blocks(bits, 1, ZERO, np.full(32, v, dtype=np.int16))
    assert xa[0] == 16 - bits
    done, bad, back, _ = oracle.decode_blocks(bits, 1, ZERO, xa, 1, 64)
    assert (back == v).all()
    # full-scale square wave: representable only with the coarsest step of filter 0
    sq = np.tile(np.array([32767] * 4 + [-32768] * 4, dtype=np.int16), 4)
    xa, _ = oracle.encode_search_blocks(bits, 1, ZERO, sq)
    done, bad, back, _ = oracle.decode_blocks(bits, 1, ZERO, xa, 1, 64)
    assert np.abs(back.astype(np.int32) - sq).max() < (1 << (16 - bits))


def test_oracle_search_in_pieces_equals_one_go(oracle):
    bits, channels, frames = 4, 2, 32 * 40
    pcm = synth.make_pcm(5, 77, channels, frames)
    whole, state = oracle.encode_search_blocks(bits, channels, ZERO, pcm)
    parts, st = [], ZERO
    for a, b in ((0, 32 * 7), (32 * 7, 32 * 8), (32 * 8, frames)):
        xa, st = oracle.encode_search_blocks(bits, channels, st, pcm[a * channels:b * channels])
        parts.append(xa)
    assert np.array_equal(np.concatenate(parts), whole) and st == state


def test_product_arithmetic_equals_the_oracle(oracle):
    """xa_core.h's search_sample / put_code, stepped by the emulator."""
    from emul_binding import Emul
    emul = Emul()
    specs = []
    for i, (bits, ch) in enumerate([(4, 1), (4, 2), (6, 1), (6, 2), (8, 1), (8, 2)] * 2):
        specs.append(dict(bits=bits, channels=ch, frames=32 * (3 + 5 * i) + (i * 11) % 32, key=300 + i))
    descs, arena, xa_bytes, pcms = batchgen.encode_batch(specs, xa_gap=3)
    rng = np.random.default_rng(3)
    descs["prev"] = rng.integers(-20000, 20000, size=(len(specs), 2, 2), dtype=np.int16)
    rc, dst, prev = emul.search(descs, arena, xa_bytes + 16)
    assert rc == 0
    end = 0
    for i, s in enumerate(specs):
        want, st = oracle.encode_search_blocks(s["bits"], s["channels"],
                                               descs[i]["prev"].tolist(), pcms[i])
        off = int(descs[i]["xa_off"])
        assert np.array_equal(dst[off:off + want.size], want), (i, s)
        assert (dst[end:off] == 0xCD).all()
        end = off + want.size
        for c in range(s["channels"]):
            assert prev[i, c].tolist() == st[c], (i, s, c)
    assert (dst[end:] == 0xCD).all()
