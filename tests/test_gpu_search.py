"""GPU tests of the searching encoder (BJXA_PLAN_ENCODE_SEARCH, an extension: the
reference has no search, src/libbjxa.c:679).  The authority for the bytes is the
plain-C restatement in oracle/ (itself checked block by block against the pinned
decoder, tests/test_search.py); the properties are checked through OUR decoder
on the GPU and the unmodified reference decoder on the CPU."""
import numpy as np
import pytest

import batchgen
from bjxa_b200 import synth
from bjxa_b200.api import PLAN_DECODE, PLAN_ENCODE, PLAN_ENCODE_SEARCH

pytestmark = pytest.mark.gpu


def run_plan(lib, kind, descs, arena, dst_bytes):
    d_src = lib.gpu_alloc(max(arena.size, 16))
    d_dst = lib.gpu_alloc(dst_bytes + 64)
    try:
        lib.upload(d_src, arena)
        lib.upload(d_dst, np.full(dst_bytes + 64, 0xCD, dtype=np.uint8))
        plan = lib.plan_create(kind, descs)
        lib.plan_run(plan, d_dst, dst_bytes + 64, d_src, arena.size)
        out = lib.plan_fetch(plan, descs.size)
        launches = lib.plan_launches(plan)
        lib.plan_free(plan)
        dst = lib.download(d_dst, dst_bytes + 64)
    finally:
        lib.gpu_free(d_src)
        lib.gpu_free(d_dst)
    return out, dst, launches


def mixed_specs():
    specs, k = [], 0
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for frames in (1, 31, 32, 33, 1000, 32 * 70 + 5, 32 * 300):
                k += 1
                specs.append(dict(bits=bits, channels=ch, frames=frames, key=500 + k))
    return specs


def test_search_equals_oracle_mixed_batch(lib, oracle):
    specs = mixed_specs()
    descs, arena, xa_bytes, pcms = batchgen.encode_batch(specs, xa_gap=5)
    rng = np.random.default_rng(9)
    descs["prev"] = rng.integers(-30000, 30000, size=(len(specs), 2, 2), dtype=np.int16)
    out, xa, launches = run_plan(lib, PLAN_ENCODE_SEARCH, descs, arena, xa_bytes)
    assert launches == 6 and (out["result"] == out["blocks"]).all() and (out["error"] == 0).all()
    end = 0
    for i, s in enumerate(specs):
        want, st = oracle.encode_search_blocks(s["bits"], s["channels"],
                                               descs[i]["prev"].tolist(), pcms[i])
        off = int(descs[i]["xa_off"])
        assert np.array_equal(xa[off:off + want.size], want), (i, s)
        assert (xa[end:off] == 0xCD).all(), (i, s, "wrote into the gap")
        end = off + want.size
        for c in range(s["channels"]):
            assert out[i]["prev"][c].tolist() == st[c], (i, s, c)
    assert (xa[end:] == 0xCD).all()


def test_search_in_pieces_equals_one_go(lib, oracle):
    """A stream encoded in three calls, the state carried through `prev`, gives
    the bytes of one call."""
    bits, ch, frames = 6, 2, 32 * 64
    pcm = synth.make_pcm(8, 4242, ch, frames)
    want, _ = oracle.encode_search_blocks(bits, ch, [[0, 0], [0, 0]], pcm)
    got, prev = [], np.zeros((2, 2), dtype=np.int16)
    for a, b in ((0, 5), (5, 6), (6, 64)):
        part = np.ascontiguousarray(pcm[a * 32 * ch:b * 32 * ch])
        descs = batchgen.make_descs(1)
        descs[0]["blocks"], descs[0]["pcm_len"] = b - a, part.size * 2
        descs[0]["bits"], descs[0]["channels"], descs[0]["prev"] = bits, ch, prev
        nbytes = (b - a) * ch * (4 * bits + 1)
        out, xa, _ = run_plan(lib, PLAN_ENCODE_SEARCH, descs, part.view(np.uint8), nbytes)
        got.append(xa[:nbytes])
        prev = out[0]["prev"].copy()
    assert np.array_equal(np.concatenate(got), want)


def test_search_round_trip_never_worse_than_plain(lib, ref):
    """encode (search) -> decode on the GPU vs encode (reference-exact) -> decode:
    per block-channel the search's squared error is never larger; the searched
    stream also decodes with the unmodified reference decoder to the same PCM."""
    specs = [dict(bits=b, channels=c, frames=32 * 500 + 9, key=900 + 10 * b + c)
             for b in (4, 6, 8) for c in (1, 2)]
    descs, arena, xa_bytes, pcms = batchgen.encode_batch(specs)
    errs = {}
    for kind in (PLAN_ENCODE, PLAN_ENCODE_SEARCH):
        out, xa, _ = run_plan(lib, kind, descs, arena, xa_bytes)
        # decode what was just encoded: the XA arena becomes the source, layout unchanged
        dd = descs.copy()
        dd["prev"] = 0
        pcm_total = int(dd["pcm_off"][-1] + dd["blocks"][-1] * 64 * dd["channels"][-1])
        o2, pcm, _ = run_plan(lib, PLAN_DECODE, dd, np.ascontiguousarray(xa[:xa_bytes]), pcm_total)
        assert (o2["result"] == o2["blocks"]).all()
        per = []
        for i, s in enumerate(specs):
            off, n = int(dd[i]["pcm_off"]), pcms[i].size
            back = pcm[off:off + 2 * n].view(np.int16)
            d = pcms[i].astype(np.int64) - back.astype(np.int64)
            frames = n // s["channels"]
            e = np.zeros((dd[i]["blocks"] * 32, s["channels"]), dtype=np.int64)
            e[:frames] = (d * d).reshape(-1, s["channels"])
            per.append(e.reshape(-1, 32, s["channels"]).sum(axis=1))
            if kind == PLAN_ENCODE_SEARCH:
                xoff = int(dd[i]["xa_off"])
                xbytes = int(dd[i]["blocks"]) * s["channels"] * (4 * s["bits"] + 1)
                hdr = synth.xa_header(xbytes, frames, 44100, s["bits"], s["channels"])
                wav = ref.xa_to_wav(hdr + xa[xoff:xoff + xbytes].tobytes())
                assert np.array_equal(np.frombuffer(wav[44:], dtype=np.int16), back), (i, s)
        errs[kind] = per
    for a, b in zip(errs[PLAN_ENCODE_SEARCH], errs[PLAN_ENCODE]):
        assert (a <= b).all()
        assert a.sum() * 4 < b.sum()


def test_search_many_streams(lib, oracle):
    """More stream-channels than one wave of warps; oracle check on a sample."""
    n = 3000
    specs = [dict(bits=(4, 6, 8)[i % 3], channels=1 + i % 2, frames=32 * (2 + i % 9) + i % 32,
                  key=7000 + i) for i in range(n)]
    descs, arena, xa_bytes, pcms = batchgen.encode_batch(specs)
    out, xa, _ = run_plan(lib, PLAN_ENCODE_SEARCH, descs, arena, xa_bytes)
    assert (out["result"] == out["blocks"]).all()
    for i in range(0, n, 37):
        s = specs[i]
        want, st = oracle.encode_search_blocks(s["bits"], s["channels"], [[0, 0], [0, 0]], pcms[i])
        off = int(descs[i]["xa_off"])
        assert np.array_equal(xa[off:off + want.size], want), (i, s)
