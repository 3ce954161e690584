"""GPU parity tests, encode side (reference-exact mode: profile byte 0 and
top-bits truncation, /root/reference/src/libbjxa.c:679,349-391).  The
reference's own tests pin no encode output (test/test_bjxa.sh:69-88 only checks
argument errors), so the authorities are the oracle, the derived goldens
generated from the compiled reference, and the compiled reference itself."""
import errno

import numpy as np
import pytest

import batchgen
from bjxa_b200 import synth
from bjxa_b200.api import PLAN_ENCODE
from conftest import sha1

pytestmark = pytest.mark.gpu


def run_plan(lib, descs, arena, xa_bytes):
    d_src = lib.gpu_alloc(max(arena.size, 16))
    d_dst = lib.gpu_alloc(xa_bytes + 64)
    try:
        lib.upload(d_src, arena)
        lib.upload(d_dst, np.full(xa_bytes + 64, 0xCD, dtype=np.uint8))
        plan = lib.plan_create(PLAN_ENCODE, descs)
        lib.plan_run(plan, d_dst, xa_bytes + 64, d_src, arena.size)
        out = lib.plan_fetch(plan, descs.size)
        lib.plan_free(plan)
        xa = lib.download(d_dst, xa_bytes + 64)
    finally:
        lib.gpu_free(d_src)
        lib.gpu_free(d_dst)
    return out, xa


def test_derived_goldens_whole_files(lib, vectors, golden):
    """Reference `bjxa encode` of the two source WAVs (hashes produced by
    running the compiled reference: tests/golden/make_golden.py)."""
    for name in ("square-mono.wav", "square-stereo.wav"):
        for bits in (4, 6, 8):
            xa = lib.wav_to_xa(vectors[name], bits)
            assert sha1(xa) == golden["derived"]["encode_sha1"][f"{name}:{bits}"]


def test_differential_goldens(lib, golden):
    seed = golden["differential"]["seed"]
    for c in golden["differential"]["encode"]:
        pcm = synth.make_pcm(seed, c["key"], c["channels"], c["frames"])
        wav = synth.riff_header(pcm.size * 2, c["channels"]) + pcm.tobytes()
        assert sha1(lib.wav_to_xa(wav, c["bits"])) == c["xa_sha1"], c


def test_plan_mixed_batch(lib, oracle):
    specs, k = [], 0
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for frames in (1, 2, 31, 32, 33, 1000, 32 * 256, 32 * 256 + 1,
                           32 * (2 * 256 + 3) + 17, 100003):
                k += 1
                specs.append(dict(bits=bits, channels=ch, frames=frames, key=k))
    descs, arena, xa_bytes, pcms = batchgen.encode_batch(specs, xa_gap=7)
    out, xa = run_plan(lib, descs, arena, xa_bytes)
    assert (out["result"] == out["blocks"]).all()
    batchgen.check_encode(oracle, specs, descs, pcms, xa[:xa_bytes])
    assert (xa[xa_bytes:] == 0xCD).all()


def test_block_at_a_time_like_the_cli(lib, ref):
    """src/bjxa_encode.c:108-169: one bjxa_encode per block, short last block."""
    pcm = synth.make_pcm(41, 1, 2, 32 * 40 + 13)
    wav = synth.riff_header(pcm.size * 2, 2) + pcm.tobytes()
    want = ref.wav_to_xa(wav, 6)
    enc = lib.encoder()
    rc, fmt = lib.parse_riff_header(wav[:44])
    assert lib.encode_init(enc, fmt, 6) == 0
    _, fmt = lib.encode_format(enc)
    raw = pcm.view(np.uint8)
    out = bytearray()
    blk = np.zeros(fmt.block_size_xa, dtype=np.uint8)
    for b in range(fmt.blocks):
        chunk = np.zeros(fmt.block_size_pcm, dtype=np.uint8)
        piece = raw[b * fmt.block_size_pcm:(b + 1) * fmt.block_size_pcm]
        chunk[:piece.size] = piece
        chunk[piece.size:] = 0x55            # garbage past the stream's end
        assert lib.encode(enc, blk, blk.size, chunk, chunk.size) == 1
        out += blk.tobytes()
    assert bytes(out) == want[32:]
    assert lib.encode(enc, blk, blk.size, chunk, chunk.size) == -1
    assert lib.errno() == errno.EPROTO
    lib.free_encoder(enc)


def test_batch_encode_host_api(lib, oracle):
    specs = [dict(bits=(4, 6, 8)[i % 3], channels=1 + i % 2, frames=500 + 91 * i, key=i)
             for i in range(30)]
    encs, dsts, srcs, pcms = [], [], [], []
    for s in specs:
        pcm = synth.make_pcm(51, s["key"], s["channels"], s["frames"])
        enc = lib.encoder()
        rc, fmt = lib.parse_riff_header(synth.riff_header(pcm.size * 2, s["channels"]))
        assert lib.encode_init(enc, fmt, s["bits"]) == 0
        encs.append(enc)
        dsts.append(np.full(fmt.blocks * fmt.block_size_xa, 0xEE, dtype=np.uint8))
        buf = np.zeros(max(pcm.size * 2, fmt.block_size_pcm), dtype=np.uint8)
        buf[:pcm.size * 2] = pcm.view(np.uint8)
        srcs.append(buf)
        pcms.append(pcm)
    dsts[3] = np.zeros(5, dtype=np.uint8)
    res, errs = lib.batch_encode(encs, dsts, srcs)
    for i, s in enumerate(specs):
        if i == 3:
            assert (res[i], errs[i]) == (-1, errno.ENOBUFS)
            continue
        want = oracle.encode_blocks(s["bits"], s["channels"], pcms[i])
        assert res[i] == (s["frames"] + 31) // 32 and errs[i] == 0
        assert np.array_equal(dsts[i], want), i
    for e in encs:
        lib.free_encoder(e)


def test_encode_then_decode_roundtrip_large(lib):
    """encode -> decode keeps exactly the top `bits` bits of every sample
    (profile 0): a size-independent property checked on ~50 M samples."""
    from bjxa_b200.api import PLAN_DECODE, make_descs
    n, frames, ch, bits = 64, 32 * 12000 + 5, 2, 4
    blocks = (frames + 31) // 32
    pcm = synth.make_pcm(61, 1, ch, frames)
    raw = pcm.view(np.uint8)
    pitch_pcm = batchgen.align16(blocks * 64 * ch)
    pitch_xa = blocks * ch * synth.block_size(bits)
    descs = make_descs(n)
    for i in range(n):
        descs[i]["xa_off"] = i * pitch_xa
        descs[i]["pcm_off"] = i * pitch_pcm
        descs[i]["blocks"] = blocks
        descs[i]["pcm_len"] = raw.size
        descs[i]["bits"], descs[i]["channels"] = bits, ch
    d_pcm = lib.gpu_alloc(n * pitch_pcm)
    d_xa = lib.gpu_alloc(n * pitch_xa + 16)
    d_back = lib.gpu_alloc(n * pitch_pcm)
    row = np.zeros(pitch_pcm, dtype=np.uint8)
    row[:raw.size] = raw
    lib.upload(d_pcm, np.tile(row, n))
    pe = lib.plan_create(PLAN_ENCODE, descs)
    lib.plan_run(pe, d_xa, n * pitch_xa + 16, d_pcm, n * pitch_pcm)
    lib.plan_fetch(pe, n)
    pd = lib.plan_create(PLAN_DECODE, descs)
    lib.plan_run(pd, d_back, n * pitch_pcm, d_xa, n * pitch_xa + 16)
    out = lib.plan_fetch(pd, n)
    assert (out["result"] == blocks).all()
    back = lib.download(d_back, n * pitch_pcm).reshape(n, pitch_pcm)
    want = (pcm & np.int16(-(1 << (16 - bits)))).view(np.uint8)
    assert (back[:, :raw.size] == want).all()
    for p in (pe, pd):
        lib.plan_free(p)
    for d in (d_pcm, d_xa, d_back):
        lib.gpu_free(d)
