"""The tile code (bjxa_b200/csrc/xa_tile.h, xa_plan.h) single-stepped on the
CPU and compared with the oracle.  This exercises the exact phase code the
sm_100a kernels run -- scheduling, carries across tiles, swizzled staging,
truncation, misaligned sources -- without a GPU.  The harness
(tests/emul/xa_emul.cc) is test-only; the product has no CPU path."""
import numpy as np
import pytest

import batchgen
from bjxa_b200 import synth
from emul_binding import Emul


@pytest.fixture(scope="module", params=["stereo-staged", "stereo-direct", "pooled", "split", "relay"])
def emul(request):
    """Every decode case runs three times: stereo streams through the staged form,
    through the direct form of the tile code (mono always uses the direct one), and
    both direct forms with their chains walked the way the pooled kernel does it
    (walk_begin / walk_block, units on 64 threads); and long-strip tiles through the
    split form: pass 1 lists the heads, pass 2 (xa_walk.h) walks every chain to its
    end across tile boundaries; and through the relay form: the tiles walk their
    own chains but hand the stragglers and everything that crosses a strip's end to
    pass 2."""
    e = Emul()
    e.stereo_direct(0 if request.param == "stereo-staged" else 1)
    e.pool(1 if request.param == "pooled" else 0)
    e.split(1 if request.param == "split" else 0)
    e.relay(1 if request.param == "relay" else 0)
    yield e
    e.stereo_direct(0)
    e.pool(0)
    e.split(0)
    e.relay(0)


STRIP_MODES = [1, 32]      # one long strip per tile / 32 short ones


def _run_decode(emul, oracle, specs, order=0, strips=0, **kw):
    descs, arena, pcm_bytes, pays = batchgen.decode_batch(specs, **kw)
    rc, dst, prev, bad = emul.decode(descs, arena, pcm_bytes + 64, order, strips)
    assert rc == 0
    return batchgen.check_decode(oracle, specs, descs, pays, dst, prev, bad)


@pytest.mark.parametrize("bits", [4, 6, 8])
@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("mix", ["P0", "P1", "P2", "P3"])
@pytest.mark.parametrize("strips", STRIP_MODES)
def test_decode_one_stream_multi_tile(emul, oracle, bits, ch, mix, strips):
    tb = emul.dec_tile_blocks(ch)
    samples = 32 * (2 * tb + 37) + 11          # 3 long tiles, ragged last block
    specs = [dict(bits=bits, channels=ch, samples=samples, mix=mix,
                  prev=((1234, -4321), (-77, 31000)), key=bits * 10 + ch)]
    _run_decode(emul, oracle, specs, strips=strips)


@pytest.mark.parametrize("mix", ["P1", "P3"])
def test_decode_many_streams_wide(emul, oracle, mix):
    """Wide tiles over many streams of different lengths, so the last tiles of
    every step are only partly filled."""
    specs = [dict(bits=(8, 4, 6)[i % 3], channels=1 + (i // 3) % 2,
                  samples=32 * (20 + (i * 7) % 50) + i % 32, mix=mix, key=700 + i,
                  prev=((i, -i), (3 * i, 1))) for i in range(240)]
    _run_decode(emul, oracle, specs, strips=32, xa_gap=4)


@pytest.mark.parametrize("alt", [0, 1])
def test_decode_class_with_both_tile_lists(emul, oracle, alt):
    """A class large enough to get both tile shapes: the carry slots are sized for
    the finer shape and must serve either list."""
    specs = [dict(bits=8 if i % 5 else 4, channels=1, samples=32 * (1 + (i * 13) % 90 + (1100 if i % 97 == 0 else 0)) + i % 32,
                  mix=("P3", "P2", "P1")[i % 3], key=4000 + i, prev=((i, -i), (0, 0)))
             for i in range(1400)]
    emul.use_alt(alt)
    try:
        _run_decode(emul, oracle, specs, strips=0, xa_gap=1)
    finally:
        emul.use_alt(0)


@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("strips", STRIP_MODES)
def test_decode_mixed_batch(emul, oracle, order, strips):
    specs = []
    k = 0
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for samples in (1, 31, 32, 33, 700, 32 * 600 + 5, 32 * 1500):
                k += 1
                specs.append(dict(bits=bits, channels=ch, samples=samples,
                                  mix=synth.MIXES[k % 4], key=100 + k,
                                  prev=((k, -k), (7 * k, 3))))
    _run_decode(emul, oracle, specs, order=order, strips=strips, xa_gap=5)


@pytest.mark.parametrize("strips", STRIP_MODES)
def test_decode_every_tail_length(emul, oracle, strips):
    specs = [dict(bits=(4, 6, 8)[r % 3], channels=1 + r % 2, samples=64 + r,
                  mix="P2", key=300 + r) for r in range(32)]
    _run_decode(emul, oracle, specs, strips=strips, xa_gap=3)


@pytest.mark.parametrize("strips", STRIP_MODES)
def test_decode_exact_tile_multiples(emul, oracle, strips):
    specs = []
    for ch in (1, 2):
        tb = emul.strip_blocks(strips, ch)
        for nt in (1, 2):
            specs.append(dict(bits=8, channels=ch, samples=32 * tb * nt, mix="P3",
                              key=400 + ch * 10 + nt))
            specs.append(dict(bits=4, channels=ch, samples=32 * tb * nt + 1, mix="P2",
                              key=450 + ch * 10 + nt))
    _run_decode(emul, oracle, specs, strips=strips)


@pytest.mark.parametrize("strips", STRIP_MODES)
def test_decode_bad_profile(emul, oracle, strips):
    """filter >= 5: the stream reports its first bad block; everything before
    it is exact (src/libbjxa.c:550,634,642)."""
    tb = emul.dec_tile_blocks(2)
    specs = [
        dict(bits=8, channels=1, samples=32 * 40, mix="P2", key=1, patch={17: 0xFF}),
        dict(bits=6, channels=2, samples=32 * 40, mix="P2", key=2, patch={2 * 9 + 1: 0x50}),
        dict(bits=4, channels=2, samples=32 * (tb + 40), mix="P3", key=3,
             patch={2 * (tb + 3): 0x9A, 2 * (tb + 20) + 1: 0xF0}),
        dict(bits=8, channels=2, samples=32 * 10, mix="P1", key=4, patch={0: 0x77}),
        dict(bits=8, channels=1, samples=32 * 10, mix="P0", key=5),
    ]
    descs, arena, pcm_bytes, pays = batchgen.decode_batch(specs)
    rc, dst, prev, bad = emul.decode(descs, arena, pcm_bytes + 64, 0, strips)
    assert rc == 0
    assert list(bad) == [17, 19, 2 * (tb + 3), 0, 0xFFFFFFFF]
    batchgen.check_decode(oracle, specs, descs, pays, dst, prev, bad)


@pytest.mark.parametrize("bits", [4, 6, 8])
@pytest.mark.parametrize("ch", [1, 2])
def test_decode_predictor_extremes(emul, oracle, bits, ch):
    """Both rails, every filter, every sign of the truncating division: the
    walkers' biased predictor step against the reference's arithmetic."""
    n = _run_decode(emul, oracle, batchgen.extremes(bits, ch))
    assert n > 72 * 32 * 39 * ch


def test_decode_saturation_vector(emul, oracle):
    """/root/reference/test/test_decode.sh:80-122 through the tile code."""
    pay = np.frombuffer(b"\x20" + b"\x7f" * 32 + b"\x20" + b"\x80" * 32, dtype=np.uint8)
    descs = batchgen.make_descs(1)
    descs[0]["blocks"] = 1
    descs[0]["pcm_len"] = 128
    descs[0]["bits"] = 8
    descs[0]["channels"] = 2
    rc, dst, prev, bad = emul.decode(descs, pay, 128)
    pcm = dst[:128].view(np.int16).reshape(32, 2)
    assert pcm[0, 0] == 32512 and (pcm[1:, 0] == 32767).all() and (pcm[:, 1] == -32768).all()


@pytest.mark.parametrize("strips", STRIP_MODES)
def test_decode_chunked_equals_whole(emul, oracle, strips):
    """Feeding a stream in pieces with the state carried by the caller gives the
    same bytes (the codec object is the checkpoint: libbjxa.c:570-571,654-655)."""
    bits, ch, blocks = 6, 2, 900
    pay = synth.xa_payload(9, 9, bits, ch, blocks, "P3")
    bs = synth.block_size(bits) * ch
    whole = oracle.decode_blocks(bits, ch, [[5, 6], [7, 8]], pay, blocks, blocks * 64 * ch)
    state = np.array([[5, 6], [7, 8]], dtype=np.int16)
    out = []
    cuts = [0, 1, 2, 300, 301, 812, 900]
    for a, b in zip(cuts[:-1], cuts[1:]):
        d = batchgen.make_descs(1)
        d[0]["blocks"] = b - a
        d[0]["pcm_len"] = (b - a) * 64 * ch
        d[0]["bits"], d[0]["channels"] = bits, ch
        d[0]["prev"] = state
        rc, dst, prev, bad = emul.decode(d, pay[a * bs:b * bs], (b - a) * 64 * ch, 0, strips)
        out.append(dst[:(b - a) * 64 * ch])
        state = prev[0]
    assert np.array_equal(np.concatenate(out).view(np.int16), whole[2])
    assert np.array_equal(state, whole[3])


@pytest.mark.parametrize("strips", STRIP_MODES)
def test_decode_arena_tail_not_multiple_of_16(emul, oracle, strips):
    """The last stream ends flush with an arena whose size is not a multiple of
    16: the bulk copy must stop short and the tail be fetched bytewise."""
    for trim in range(0, 16):
        specs = [dict(bits=8, channels=1, samples=32 * 3 + trim, mix="P2", key=trim)]
        descs, arena, pcm_bytes, pays = batchgen.decode_batch(specs)
        pre = np.zeros(trim, dtype=np.uint8)
        descs[0]["xa_off"] = trim
        rc, dst, prev, bad = emul.decode(descs, np.concatenate([pre, arena]), pcm_bytes + 64,
                                         0, strips)
        batchgen.check_decode(oracle, specs, descs, pays, dst, prev, bad)


@pytest.mark.parametrize("order", [0, 2])
def test_encode_mixed_batch(emul, oracle, order):
    te = emul.enc_tile_blocks()
    specs = []
    k = 0
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for frames in (1, 2, 31, 32, 33, 1000, 32 * te, 32 * te + 1, 32 * (2 * te + 3) + 17):
                k += 1
                specs.append(dict(bits=bits, channels=ch, frames=frames, key=k))
    descs, arena, xa_bytes, pcms = batchgen.encode_batch(specs, xa_gap=7)
    rc, dst = emul.encode(descs, arena, xa_bytes, order)
    assert rc == 0
    batchgen.check_encode(oracle, specs, descs, pcms, dst)


def test_plan_order_and_validation(emul):
    """Host planning logic: every strip of a stream after its predecessor,
    time-major issue, the tile shape follows the stream count, EINVAL on bad
    descriptors."""
    tb1, tb2 = emul.strip_blocks(1, 1), emul.strip_blocks(1, 2)
    d = batchgen.make_descs(5)
    for i, (bits, ch, blocks) in enumerate([(8, 1, 3 * tb1 + 1), (8, 1, tb1), (4, 2, 2 * tb2),
                                            (8, 1, 0), (8, 1, 2 * tb1 + 5)]):
        d[i]["bits"], d[i]["channels"], d[i]["blocks"] = bits, ch, blocks
        d[i]["pcm_len"] = blocks * 64 * ch
    n, ts, tc, tj, tb, slots, nsb = emul.plan(0, d)
    assert n == 4 + 1 + 2 + 0 + 3 and slots == n and (tc == 1).all()
    assert list(nsb) == [1] * 6             # few streams: one long strip per tile
    b81 = slice(tb[4], tb[5])
    assert list(zip(ts[b81], tj[b81])) == [(0, 0), (4, 0), (1, 0), (0, 1), (4, 1),
                                           (0, 2), (4, 2), (0, 3)]
    assert list(zip(ts[tb[1]:tb[2]], tj[tb[1]:tb[2]])) == [(2, 0), (2, 1)]
    for field, val in (("bits", 5), ("channels", 3), ("pcm_off", 8), ("pcm_len", 3)):
        e = d.copy()
        e[0][field] = val
        assert emul.plan(0, e)[0] == -22

    # wide tiles of `wide` strips, partly filled at the end of a step
    w = emul.wide
    sb = emul.strip_blocks(w, 1)
    m = 2 * w + 5
    d = batchgen.make_descs(m)
    d["bits"], d["channels"] = 8, 1
    d["blocks"] = [3 * sb if i < w + 2 else sb + 1 for i in range(m)]
    d["pcm_len"] = d["blocks"] * 64
    assert emul.plan(0, d)[6][4] == 1       # too few streams for the automatic choice
    n, ts, tc, tj, tb, slots, nsb = emul.plan(0, d, strips=w)
    assert nsb[4] == w
    assert list(tc) == [w, w, 5, w, w, 5, w, 2]
    assert list(tj) == [0, 0, 0, 1, 1, 1, 2, 2]
    assert slots == (w + 2) * 3 + (w + 3) * 2
    big = batchgen.make_descs(20000)
    big["bits"], big["channels"], big["blocks"], big["pcm_len"] = 4, 2, 3, 3 * 128
    # a large class gets both shapes (the census picks at launch): the long-strip
    # list first, the wide one after every primary list; slots fit the finer shape
    n, ts, tc, tj, tb, slots, nsb = emul.plan(0, big)
    ab, an = emul.alt
    assert nsb[1] == 1 and an[1] == w and list(an[[0, 2, 3, 4, 5]]) == [0] * 5
    assert tb[2] - tb[1] == 20000 and ab[1] == tb[6] and ab[2] == n
    sb2 = emul.strip_blocks(w, 2)
    assert (tc[ab[1]:ab[2]] == w).all() and ab[2] - ab[1] == (20000 // w) * -(-3 // sb2)
    assert slots == 20000 * -(-3 // sb2)


# ---- the segment form (xa_walk.h: walk_seg_serial, the semantics of xa_seg_kernel) ----

@pytest.fixture()
def seg_emul():
    e = Emul()
    e.seg(1)
    yield e
    e.seg(0)


@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("bits,ch", [(4, 1), (6, 1), (8, 1), (4, 2), (6, 2), (8, 2)])
def test_seg_form_class(seg_emul, oracle, bits, ch, order):
    """One class, lengths from one item to several segments (ragged last blocks), so
    that a tile's 32 lanes belong to many streams; every mix: P0/P1 find a cut block
    right in front of a segment, P2 a few items back, P3 never -- those lanes wait
    for the lane in front (another pass over the tile) or, lane 0, for the mailbox of
    the tile before; streams shorter than the look-back start from their own state."""
    e = seg_emul
    seg, n = e.seg_items, 73
    specs = [dict(bits=bits, channels=ch,
                  samples=32 * (1 + (i * 37) % (3 * seg + 5)) - (i % 32 if i % 3 else 0),
                  mix=("P2", "P3", "P1", "P0", "P2")[i % 5], key=9000 + 100 * bits + 10 * ch + i,
                  prev=((i, -i), (7 * i, -3))) for i in range(n)]
    specs[0]["samples"] = 32 * (4 * seg + 3) + 5        # the longest: five segments
    specs[1]["samples"] = 32 * (4 * seg) + 32           # P3, mailbox hand-overs all the way
    _run_decode(e, oracle, specs, order=order, xa_gap=3)


@pytest.mark.parametrize("bits,ch", [(4, 1), (8, 1), (6, 2), (8, 2)])
def test_seg_form_look_back_limits(seg_emul, oracle, bits, ch):
    """Chains that reach just inside and just outside the look-back in front of a
    segment's first item, per channel: a cut block exactly kSegBack items back is
    found, one item further is not (the state then comes through the mailbox); an
    invalid profile counts as a cut and is reported once, by the lane whose segment
    holds it."""
    e = seg_emul
    seg, back, n = e.seg_items, e.seg_back, 64
    specs = []
    for i in range(n):
        blocks = 2 * seg + 17
        patch = {}
        # all chain blocks (mix P3) but for the patched ones; item m, channel c = m * ch + c
        m0 = seg - back - 1 + (i % 3)                    # back + 1, back, back - 1 items in front
        patch[m0 * ch] = 0x03                            # channel 0: a cut block, range 3
        if ch == 2:
            patch[(seg - 1 - (i % 5)) * ch + 1] = 0x00   # channel 1: a cut close by
            if i % 4 == 0:
                patch[(2 * seg - 1 - back - (i % 8)) * ch + 1] = 0x0f
        if i % 7 == 0:
            patch[(seg + 5) * ch + (ch - 1)] = 0x5a      # invalid filter inside segment 1
        if i % 11 == 0:
            patch[(seg - 2) * ch] = 0xf1                 # ... and as the "cut" in front of it
        specs.append(dict(bits=bits, channels=ch, samples=32 * blocks - 3 * (i % 9),
                          mix="P3", key=9900 + i, patch=patch, prev=((11, -12), (13, -14))))
    _run_decode(e, oracle, specs, xa_gap=1)


@pytest.mark.parametrize("order", [0, 2])
@pytest.mark.parametrize("bits,ch", [(8, 1), (4, 2), (6, 1)])
def test_seg_form_long_streams(seg_emul, oracle, bits, ch, order):
    """Streams that fill several tiles by themselves (32 lanes = 32 consecutive
    segments of one stream) next to short ones that share tiles.  P2: every lane
    finds its state a few items back; P3: every lane but the first waits for the lane
    in front of it -- 32 passes over the tile -- and the first for the mailbox of the
    tile before; a P3 stream with a few cut blocks: some lanes wait, some do not."""
    e = seg_emul
    seg, long_items = e.seg_items, 64 * e.seg_items
    specs = [
        dict(bits=bits, channels=ch, samples=32 * (long_items + 3 * seg + 7) - 5, mix="P2", key=31000,
             prev=((100, -100), (7, 8))),
        dict(bits=bits, channels=ch, samples=32 * (long_items + 1), mix="P3", key=31001,
             prev=((-5, 6), (70, -80))),
        dict(bits=bits, channels=ch, samples=32 * (long_items + 40 * seg) - 31, mix="P3", key=31002,
             patch={(5 * seg - 3) * ch: 0x02, (9 * seg - 50) * ch + (ch - 1): 0x0c,
                    (33 * seg + 1) * ch: 0x60, (36 * seg - 1) * ch: 0x00, (36 * seg - 1) * ch + ch - 1: 0x04}),
        dict(bits=bits, channels=ch, samples=32 * long_items, mix="P1", key=31003),
    ]
    specs += [dict(bits=bits, channels=ch, samples=32 * (1 + (i * 53) % (2 * seg)) - i % 32,
                   mix=("P2", "P3", "P0")[i % 3], key=31100 + i, prev=((i, -i), (2 * i, 3))) for i in range(40)]
    _run_decode(e, oracle, specs, order=order, xa_gap=2)
