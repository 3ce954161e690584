"""Host side of the drop-in boundary: codec objects, headers, errno contract.
Mirrors /root/reference/test/test_libbjxa_api.c (the parts that need no GPU)
and the header cases of test/test_decode_error.sh; differential against the
compiled reference where it is available.  CPU only."""
import ctypes as C
import errno

import numpy as np
import pytest

from bjxa_b200 import synth
from bjxa_b200.api import DESC_DTYPE
from bjxa_b200.capi import Format

JUNK = C.create_string_buffer(b"random junk", 64)


def junk():
    return C.addressof(JUNK)


def test_memory_management(lib):
    """test_libbjxa_api.c:40-69"""
    dec = lib.decoder()
    assert dec
    assert lib._bjxa_free_decoder(None) == -1 and lib.errno() == errno.EFAULT
    p = C.c_void_p(dec)
    assert lib._bjxa_free_decoder(C.byref(p)) == 0 and p.value is None
    assert lib._bjxa_free_decoder(C.byref(p)) == -1 and lib.errno() == errno.EFAULT
    p = C.c_void_p(junk())
    assert lib._bjxa_free_decoder(C.byref(p)) == -1 and lib.errno() == errno.EINVAL
    assert p.value == junk()
    enc = lib.encoder()
    p = C.c_void_p(enc)
    assert lib._bjxa_free_encoder(C.byref(p)) == 0 and p.value is None
    # a decoder is not an encoder
    dec = lib.decoder()
    p = C.c_void_p(dec)
    assert lib._bjxa_free_encoder(C.byref(p)) == -1 and lib.errno() == errno.EINVAL
    assert lib.free_decoder(dec) == 0


def test_header_parsing_errors(lib):
    """test_libbjxa_api.c:71-108"""
    dec = lib.decoder()
    keep = C.create_string_buffer(64)
    buf = C.addressof(keep)
    assert lib._bjxa_parse_header(None, buf, 32) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_parse_header(junk(), buf, 32) == -1 and lib.errno() == errno.EINVAL
    assert lib._bjxa_parse_header(dec, None, 32) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_parse_header(dec, buf, 0) == -1 and lib.errno() == errno.ENOBUFS
    assert lib._bjxa_parse_header(dec, buf, 31) == -1 and lib.errno() == errno.ENOBUFS
    assert lib._bjxa_fread_header(None, 1) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_fread_header(junk(), 1) == -1 and lib.errno() == errno.EINVAL
    assert lib._bjxa_fread_header(dec, None) == -1 and lib.errno() == errno.EFAULT
    lib.free_decoder(dec)


def test_format_errors(lib):
    """test_libbjxa_api.c:110-137"""
    dec = lib.decoder()
    fmt = Format()
    assert lib._bjxa_decode_format(None, C.byref(fmt)) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_decode_format(junk(), C.byref(fmt)) == -1 and lib.errno() == errno.EINVAL
    assert lib._bjxa_decode_format(dec, C.byref(fmt)) == -1 and lib.errno() == errno.EINVAL
    assert lib._bjxa_decode_format(dec, None) == -1 and lib.errno() == errno.EFAULT
    lib.free_decoder(dec)


def test_decode_argument_errors(lib, vectors):
    """test_libbjxa_api.c:139-176 -- everything that fails before any block is
    touched, in the reference's check order (src/libbjxa.c:612-620)."""
    src = np.zeros(4096, dtype=np.uint8)
    dst = np.zeros(4096, dtype=np.uint8)
    dec = lib.decoder()
    assert lib.decode(None, dst, 4096, src, 4096) == -1 and lib.errno() == errno.EFAULT
    assert lib.decode(junk(), dst, 4096, src, 4096) == -1 and lib.errno() == errno.EINVAL
    assert lib.decode(dec, dst, 4096, src, 4096) == -1 and lib.errno() == errno.EINVAL
    assert lib.parse_header(dec, vectors["square-mono-4.xa"][:32]) == 32
    assert lib.decode(dec, None, 4096, src, 4096) == -1 and lib.errno() == errno.EFAULT
    assert lib.decode(dec, dst, 0, src, 4096) == -1 and lib.errno() == errno.ENOBUFS
    assert lib.decode(dec, dst, 4096, None, 4096) == -1 and lib.errno() == errno.EFAULT
    assert lib.decode(dec, dst, 4096, src, 0) == -1 and lib.errno() == errno.ENOBUFS
    assert lib.decode(dec, dst, 63, src, 4096) == -1 and lib.errno() == errno.ENOBUFS
    assert lib.decode(dec, dst, 4096, src, 16) == -1 and lib.errno() == errno.ENOBUFS
    lib.free_decoder(dec)


def test_no_cpu_fallback(lib, vectors):
    """Without a CUDA device the hot path fails loudly; it never computes on
    the CPU."""
    if lib.gpu_count() > 0:
        pytest.skip("a GPU is present")
    xa = vectors["square-mono-4.xa"]
    dec = lib.decoder()
    assert lib.parse_header(dec, xa[:32]) == 32
    dst = np.full(4096, 0xEE, dtype=np.uint8)
    assert lib.decode(dec, dst, 4096, xa[32:32 + 17 * 8], 17 * 8) == -1
    assert lib.errno() == errno.ENODEV
    assert (dst == 0xEE).all()
    lib.free_decoder(dec)
    enc = lib.encoder()
    rc, fmt = lib.parse_riff_header(synth.riff_header(4000, 2))
    assert lib.encode_init(enc, fmt, 4) == 0
    out = np.full(4096, 0xEE, dtype=np.uint8)
    assert lib.encode(enc, out, 4096, dst, 4096) == -1 and lib.errno() == errno.ENODEV
    assert (out == 0xEE).all()
    lib.free_encoder(enc)
    with pytest.raises(OSError) as ei:
        lib.plan_create(0, np.zeros(0, dtype=DESC_DTYPE))
    assert ei.value.errno == errno.ENODEV


def test_riff_dump_errors(lib, vectors):
    """test_libbjxa_api.c:195-248"""
    dec = lib.decoder()
    keep = C.create_string_buffer(64)
    buf = C.addressof(keep)
    assert lib._bjxa_dump_riff_header(None, buf, 64) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_dump_riff_header(junk(), buf, 64) == -1 and lib.errno() == errno.EINVAL
    assert lib._bjxa_dump_riff_header(dec, buf, 64) == -1 and lib.errno() == errno.EINVAL
    assert lib._bjxa_fwrite_riff_header(None, 1) == -1 and lib.errno() == errno.EFAULT
    assert lib.parse_header(dec, vectors["square-mono-4.xa"][:32]) == 32
    assert lib._bjxa_fwrite_riff_header(dec, None) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_dump_riff_header(dec, None, 64) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_dump_riff_header(dec, buf, 0) == -1 and lib.errno() == errno.ENOBUFS
    assert lib._bjxa_dump_riff_header(dec, buf, 44) == 44
    lib.free_decoder(dec)


def test_pcm_dump(lib):
    """test_libbjxa_api.c:250-280"""
    src = np.arange(-16, 16, dtype=np.int16) * 1021
    dst = np.zeros(64, dtype=np.uint8)
    assert lib.dump_pcm(None, src, 32) == -1 and lib.errno() == errno.EFAULT
    assert lib.dump_pcm(dst, None, 32) == -1 and lib.errno() == errno.EFAULT
    assert lib.dump_pcm(dst, src, 0) == -1 and lib.errno() == errno.ENOBUFS
    assert lib.dump_pcm(dst, src, 31) == -1 and lib.errno() == errno.ENOBUFS
    assert lib.dump_pcm(dst, src, 64) == 0
    assert dst.tobytes() == src.astype("<i2").tobytes()
    assert lib._bjxa_fwrite_pcm(None, 32, 1) == -1 and lib.errno() == errno.EFAULT
    assert lib._bjxa_fwrite_pcm(src.ctypes.data, 0, 1) == -1 and lib.errno() == errno.ENOBUFS
    assert lib._bjxa_fwrite_pcm(src.ctypes.data, 31, 1) == -1 and lib.errno() == errno.ENOBUFS
    assert lib._bjxa_fwrite_pcm(src.ctypes.data, 32, None) == -1 and lib.errno() == errno.EFAULT


def _hdr(**kw):
    f = dict(data_len=66, samples=32, rate=44100, bits=8, channels=2)
    f.update(kw)
    return synth.xa_header(f["data_len"], f["samples"], f["rate"], f["bits"], f["channels"],
                           kw.get("prev", ((0, 0), (0, 0))))


BAD_HEADERS = {   # test/test_decode_error.sh:36-219
    "magic": b"KWD2" + _hdr()[4:],
    "nDataLen=0": _hdr(data_len=0),
    "nSamples=0": _hdr(samples=0),
    "too many samples": _hdr(samples=33),
    "not enough samples": _hdr(data_len=132, samples=32),
    "rate 0": _hdr(rate=0),
    "data not a block multiple": _hdr(data_len=67),
    "nBits=12": _hdr(bits=12),
    "nChannels=5": _hdr(channels=5),
    # the reference stops on this one with an assert in bjxa_decode_format
    # (libbjxa.c:596); a batch library reports the header instead
    "stereo, odd number of blocks": _hdr(bits=4, channels=2, data_len=51, samples=40),
}


@pytest.mark.parametrize("case", sorted(BAD_HEADERS))
def test_malformed_headers(lib, case):
    dec = lib.decoder()
    good = _hdr(prev=((1, 2), (3, 4)))
    assert lib.parse_header(dec, good) == 32
    assert lib.parse_header(dec, BAD_HEADERS[case]) == -1 and lib.errno() == errno.EPROTO
    # the decoder is updated atomically: still the good stream
    rc, fmt = lib.decode_format(dec)
    assert rc == 0 and fmt.blocks == 1 and fmt.channels == 2
    lib.free_decoder(dec)


def test_header_differential_vs_reference(lib, ref):
    """Random headers: same verdict, same format, same RIFF bytes as the
    compiled reference."""
    words = synth.rand_u64(5, 5, 4000)
    agree_ok = 0
    for i in range(2000):
        a, b = int(words[2 * i]), int(words[2 * i + 1])
        bits = (4, 6, 8, 5)[a & 3] if (a >> 40) & 7 == 0 else (4, 6, 8)[a % 3]
        ch = (1, 2, 2, 3)[(a >> 2) & 3] if (a >> 43) & 7 == 0 else 1 + ((a >> 2) & 1)
        blocks = 1 + ((a >> 8) % 50000)
        bs = 4 * bits + 1
        data_len = blocks * bs * min(ch, 2) + ((a >> 4) & 1) * ((a >> 5) & 7) * ((a >> 46) & 1)
        samples = max(0, 32 * blocks - ((b >> 3) % 36) + 2)
        rate = (b >> 20) & 0xFFFF if (b >> 50) & 7 else 0
        hdr = synth.xa_header(data_len & 0xFFFFFFFF, samples, rate, bits, ch,
                              (((b >> 1) & 0x7FFF, -3), (9, (b >> 7) & 0xFF)))
        out = []
        for L in (lib, ref):
            dec = L.decoder()
            L.clear_errno()
            rc = L.parse_header(dec, hdr)
            err = L.errno() if rc < 0 else 0
            riff = bytearray(44)
            fmtd = None
            if rc > 0:
                _, fmt = L.decode_format(dec)
                fmtd = fmt.as_dict()
                assert L.dump_riff_header(dec, riff) == 44
            L.free_decoder(dec)
            out.append((rc, err, fmtd, bytes(riff)))
        assert out[0] == out[1], (i, hdr.hex(), out)
        agree_ok += out[0][0] > 0
    assert agree_ok > 100


def test_encoder_setup_differential_vs_reference(lib, ref):
    words = synth.rand_u64(6, 6, 3000)
    ok = 0
    for i in range(1500):
        a, b = int(words[2 * i]), int(words[2 * i + 1])
        ch = (1, 2, 2, 3)[a & 3] if (a >> 40) & 3 == 0 else 1 + (a & 1)
        frames = (a >> 4) % 100000
        pcm_bytes = frames * 2 * min(ch, 2) + ((a >> 3) & 1) * ((a >> 44) & 1)
        rate = (b & 0x1FFFF) if (b >> 40) & 3 == 0 else 1 + (b & 0xFFFF) % 65000
        wav = synth.riff_header(pcm_bytes, ch, rate)
        bits = (4, 6, 8, 7)[(b >> 20) & 3]
        out = []
        for L in (lib, ref):
            L.clear_errno()
            rc, fmt = L.parse_riff_header(wav)
            rec = [rc, L.errno() if rc < 0 else 0]
            if rc > 0:
                enc = L.encoder()
                L.clear_errno()
                rc2 = L.encode_init(enc, fmt, bits)
                rec += [rc2, L.errno() if rc2 < 0 else 0, fmt.as_dict()]
                hdr = bytearray(32)
                if rc2 == 0:
                    _, f2 = L.encode_format(enc)
                    assert L.dump_header(enc, hdr) == 32
                    rec += [f2.as_dict(), bytes(hdr)]
                else:
                    L.clear_errno()
                    rec += [L.dump_header(enc, hdr), L.errno()]
                L.free_encoder(enc)
            out.append(rec)
        assert out[0] == out[1], (i, out)
        ok += len(out[0]) == 7 and isinstance(out[0][5], dict)
    assert ok > 100


def test_encode_argument_errors(lib):
    """src/libbjxa.c:770-778 order."""
    enc = lib.encoder()
    buf = np.zeros(4096, dtype=np.uint8)
    assert lib.encode(None, buf, 4096, buf, 4096) == -1 and lib.errno() == errno.EFAULT
    assert lib.encode(junk(), buf, 4096, buf, 4096) == -1 and lib.errno() == errno.EINVAL
    assert lib.encode(enc, buf, 4096, buf, 4096) == -1 and lib.errno() == errno.EINVAL
    rc, fmt = lib.parse_riff_header(synth.riff_header(4000, 2))
    assert lib.encode_init(enc, fmt, 5) == -1 and lib.errno() == errno.EINVAL
    assert lib.encode_init(enc, fmt, 6) == 0
    assert lib.encode(enc, None, 4096, buf, 4096) == -1 and lib.errno() == errno.EFAULT
    assert lib.encode(enc, buf, 4096, None, 4096) == -1 and lib.errno() == errno.EFAULT
    assert lib.encode(enc, buf, 49, buf, 4096) == -1 and lib.errno() == errno.ENOBUFS
    assert lib.encode(enc, buf, 4096, buf, 127) == -1 and lib.errno() == errno.ENOBUFS
    lib.free_encoder(enc)


def test_shard_range(lib):
    """Contiguous, exhaustive, balanced by bytes."""
    sizes = (synth.rand_u64(8, 8, 1000) % np.uint64(100000)).astype(np.uint64) + 1
    for world in (1, 2, 3, 4, 8):
        cover = []
        loads = []
        for r in range(world):
            first, count = lib.shard_range(1000, r, world, sizes)
            cover += list(range(first, first + count))
            loads.append(int(sizes[first:first + count].sum()))
        assert cover == list(range(1000))
        assert max(loads) - min(loads) <= 2 * int(sizes.max())
        first, count = lib.shard_range(1000, world - 1, world)
        assert first + count == 1000
    with pytest.raises(OSError):
        lib.shard_range(10, 3, 3)


def test_corpus_extent_and_argument_errors(lib, vectors, oracle):
    """bjxa_corpus_extent only reads headers, so it runs without a GPU: the size
    it reports is the layout bjxa_corpus_run uses (WAV files at offsets = 4 mod
    16); argument errors; and without a CUDA device bjxa_corpus_run refuses."""
    import ctypes as C

    from bjxa_b200.api import FILE_DTYPE
    names = ["square-stereo-8.xa", "square-mono-4.xa", "square-mono-6.xa"]
    files = [vectors[n] for n in names] + [b"KWD1" + bytes(10)]      # + one too short
    table = np.zeros(len(files), dtype=FILE_DTYPE)
    off = 3
    for i, f in enumerate(files):
        table[i]["in_off"], table[i]["in_len"] = off, len(f)
        off += len(f) + i
    arena = np.zeros(off + 16, dtype=np.uint8)
    for t, f in zip(table, files):
        arena[int(t["in_off"]):int(t["in_off"]) + len(f)] = np.frombuffer(f, dtype=np.uint8)
    need = C.c_uint64(0)

    def extent(kind, tab):
        return lib._bjxa_corpus_extent(kind, arena.ctypes.data, arena.size, tab.ctypes.data,
                                       tab.size, C.byref(need))
    assert extent(0, table) == 0
    cur = 0
    for f in files[:3]:
        cur = ((cur + 44 + 15) & ~15) - 44 + len(oracle.xa_to_wav(f))
    assert need.value == cur + 16
    swapped = table.copy()
    swapped[[0, 1]] = swapped[[1, 0]]
    assert extent(0, swapped) == -1 and lib.errno() == errno.EINVAL
    beyond = table.copy()
    beyond[3]["in_len"] = arena.size
    assert extent(0, beyond) == -1 and lib.errno() == errno.ENOBUFS
    assert extent(7, table) == -1 and lib.errno() == errno.EINVAL
    if lib.gpu_count() <= 0:
        out = np.full(int(need.value) + 64, 0xEE, dtype=np.uint8)
        rc = lib._bjxa_corpus_run(0, arena.ctypes.data, arena.size, out.ctypes.data, out.size,
                                  table.ctypes.data, table.size)
        assert rc == -1 and lib.errno() == errno.ENODEV and (out == 0xEE).all()


def test_scatter_argument_errors(lib):
    """bjxa_gpu_scatter_async checks its arguments before it touches the device."""
    buf = np.zeros(128, dtype=np.uint8)
    assert lib._bjxa_gpu_scatter_async(None, None, 44, 0, None) == 0          # nothing to do
    assert lib._bjxa_gpu_scatter_async(None, buf.ctypes.data, 44, 1, None) == -1
    assert lib.errno() == errno.EFAULT
    assert lib._bjxa_gpu_scatter_async(buf.ctypes.data, buf.ctypes.data, 57, 1, None) == -1
    assert lib.errno() == errno.EINVAL
