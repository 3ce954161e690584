"""Deterministic synthetic XA / PCM streams (host side, numpy, integer only).

Used by the tests, the golden-fixture generator and bench.py's CPU sample.  The
generator is a counter-based splitmix64 hash keyed by (seed, key, index), so a
stream's bytes depend on nothing but its parameters -- no numpy Generator
state, no libm.

Profile-byte mixes ("Config 2"):

  P0  every block filter 0, range uniform 0..8 -- what the reference encoder
      (and any re-encode) produces: /root/reference/src/libbjxa.c:679
  P1  "xa.exe-like": the histogram measured on the reference's
      test/square-*.xa vectors -- filters {0: 94.9 %, 1: 0.66 %, 2: 4.45 %,
      3: 0.005 %}, non-zero filters isolated, ranges {0,1,2,3,4,6}
  P2  filter uniform 0..4, range uniform 0..15
  P3  adversarial: filter uniform 1..4 (no state-independent block anywhere),
      range uniform 0..15
  PX  any byte value at all, including invalid filters >= 5 (error paths)
  Cnn every block independently a chain block (filter 1..4) with probability
      nn %, range uniform 0..15 (finds where one decode form overtakes another)
"""
from __future__ import annotations

import struct

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)

BLOCK_SAMPLES = 32
MIXES = ("P0", "P1", "P2", "P3")


def _splitmix(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + _GOLD
        z = (z ^ (z >> np.uint64(30))) * _C1
        z = (z ^ (z >> np.uint64(27))) * _C2
        return z ^ (z >> np.uint64(31))


def rand_u64(seed: int, key: int, n: int, start: int = 0) -> np.ndarray:
    """n 64-bit words of stream (seed, key), counters start..start+n-1."""
    with np.errstate(over="ignore"):
        base = _splitmix(np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))
        base = _splitmix(base ^ np.uint64(key & 0xFFFFFFFFFFFFFFFF))
        ctr = np.arange(start, start + n, dtype=np.uint64)
        return _splitmix(base + ctr * _GOLD)


def rand_bytes(seed: int, key: int, n: int) -> np.ndarray:
    words = rand_u64(seed, key, (n + 7) // 8)
    return words.view(np.uint8)[:n].copy()


def rand_unit(seed: int, key: int, n: int) -> np.ndarray:
    """n integers uniform in [0, 2**24) (used as fixed-point probabilities)."""
    return (rand_u64(seed, key, n) >> np.uint64(40)).astype(np.int64)


def block_size(bits: int) -> int:
    """/root/reference/src/libbjxa.c:431"""
    return 4 * bits + 1


def profile_bytes(mix: str, seed: int, key: int, n: int) -> np.ndarray:
    """n profile bytes (filter << 4 | range) for one channel's block sequence."""
    u = rand_unit(seed, key * 4 + 1, n)
    v = rand_unit(seed, key * 4 + 2, n)
    one = 1 << 24
    if mix == "P0":
        filt = np.zeros(n, dtype=np.int64)
        rng = (v * 9) >> 24
    elif mix == "P1":
        # cumulative thresholds for filters 1, 2, 3 (rest filter 0)
        t1 = int(0.0066 * one)
        t2 = t1 + int(0.0445 * one)
        t3 = t2 + int(0.00005 * one)
        filt = np.where(u < t1, 1, np.where(u < t2, 2, np.where(u < t3, 3, 0)))
        # isolate: a non-zero filter never follows a non-zero filter
        prev_nz = np.concatenate(([False], filt[:-1] != 0))
        for _ in range(4):  # resolve chains left to right
            clash = (filt != 0) & prev_nz
            if not clash.any():
                break
            filt = np.where(clash, 0, filt)
            prev_nz = np.concatenate(([False], filt[:-1] != 0))
        hist = np.array([8790, 6891, 3447, 487, 137, 920], dtype=np.int64)
        vals = np.array([0, 1, 2, 3, 4, 6], dtype=np.int64)
        cum = np.cumsum(hist) * one // hist.sum()
        rng = vals[np.searchsorted(cum, v, side="right").clip(0, 5)]
    elif mix == "P2":
        filt = (u * 5) >> 24
        rng = (v * 16) >> 24
    elif mix == "P3":
        filt = 1 + ((u * 4) >> 24)
        rng = (v * 16) >> 24
    elif mix == "PX":
        return (u & 0xFF).astype(np.uint8)
    elif mix.startswith("C"):
        # "Cnn": every block independently a chain block (filter 1..4) with
        # probability nn %, ranges uniform 0..15
        w = rand_unit(seed, key * 4 + 3, n)
        filt = np.where(u * 100 < int(mix[1:]) * one, 1 + ((w * 4) >> 24), 0)
        rng = (v * 16) >> 24
    else:
        raise ValueError(f"unknown profile mix {mix!r}")
    return ((filt << 4) | rng).astype(np.uint8)


def xa_header(data_len: int, samples: int, rate: int, bits: int, channels: int,
              prev=((0, 0), (0, 0)), loop: int = 0, pad: int = 0) -> bytes:
    """32-byte KWD1 header (/root/reference/bjxa.5.rst:63-105,
    src/libbjxa.c:410-421)."""
    return struct.pack("<4sIIHBBIhhhhI", b"KWD1", data_len, samples, rate, bits,
                       channels, loop, prev[0][0], prev[0][1], prev[1][0],
                       prev[1][1], pad)


def xa_payload(seed: int, key: int, bits: int, channels: int, blocks: int,
               mix: str) -> np.ndarray:
    """`blocks` effective blocks (L then R when stereo) of random payload."""
    bs = block_size(bits)
    nbc = blocks * channels
    pay = rand_bytes(seed, key * 4, nbc * bs).reshape(nbc, bs)
    for c in range(channels):
        pay[c::channels, 0] = profile_bytes(mix, seed, key * 2 + c, blocks)
    return pay.reshape(-1)


def make_xa(seed: int, key: int, bits: int, channels: int, samples: int,
            mix: str = "P2", prev=((0, 0), (0, 0)), rate: int = 44100) -> bytes:
    """A whole .xa image: header + ceil(samples/32) effective blocks."""
    blocks = (samples + BLOCK_SAMPLES - 1) // BLOCK_SAMPLES
    pay = xa_payload(seed, key, bits, channels, blocks, mix)
    return xa_header(pay.size, samples, rate, bits, channels, prev) + pay.tobytes()


def make_pcm(seed: int, key: int, channels: int, frames: int) -> np.ndarray:
    """Interleaved int16 PCM, integer-only: three triangle waves with seeded
    periods plus hash noise, clipped ("Config 4")."""
    par = rand_u64(seed, key * 4 + 3, 8)
    t = np.arange(frames, dtype=np.int64)
    out = np.empty((frames, channels), dtype=np.int16)
    for c in range(channels):
        acc = np.zeros(frames, dtype=np.int64)
        for k in range(3):
            period = 16 + int(par[(2 * k + c) % 8] >> np.uint64(52)) % 2000
            amp = 3000 + int(par[(2 * k + c + 1) % 8] >> np.uint64(50)) % 9000
            ph = (t + c * 7) % period
            tri = np.abs(2 * ph - period) * 2 - period  # [-period, period]
            acc += tri * amp // period
        noise = (rand_u64(seed, key * 8 + 4 + c, frames) >> np.uint64(52)).astype(np.int64) - 2048
        out[:, c] = np.clip(acc + noise, -32768, 32767).astype(np.int16)
    return out.reshape(-1)


def riff_header(pcm_bytes: int, channels: int, rate: int = 44100) -> bytes:
    """Canonical 44-byte PCM WAV header (/root/reference/src/libbjxa.c:909-922)."""
    return struct.pack("<4sI8sIHHIIHH4sI", b"RIFF", 36 + pcm_bytes, b"WAVEfmt ",
                       16, 1, channels, rate, rate * 2 * channels, 2 * channels,
                       16, b"data", pcm_bytes)


def stream_checksum(buf) -> int:
    """The sum bjxa_plan_checksum computes on the device (include/bjxa_batch.h):
    sum of word_i * ((i * G + C) | 1) mod 2^64 over the little-endian 32-bit words
    of `buf`, bytes past its end taken as zero."""
    b = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    pad = (-b.size) % 4
    if pad:
        b = np.concatenate((b, np.zeros(pad, dtype=np.uint8)))
    w = b.view("<u4").astype(np.uint64)
    with np.errstate(over="ignore"):
        i = np.arange(w.size, dtype=np.uint64)
        mult = (i * _GOLD + np.uint64(0xD1B54A32D192ED03)) | np.uint64(1)
        return int((w * mult).sum(dtype=np.uint64))
