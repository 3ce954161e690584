"""Builds batches (arenas + descriptor tables) for the parity tests and checks
a decoded / encoded batch against the oracle.  Shared by the CPU single-step
tests (test_emul.py) and the GPU parity tests (test_gpu_*.py)."""
import numpy as np

from bjxa_b200 import synth
from bjxa_b200.api import make_descs


def align16(x):
    return (x + 15) & ~15


def decode_batch(specs, seed=1, xa_gap=0, pcm_pad=0):
    """specs: list of dict(bits, channels, samples, mix, prev=None, key).
    Returns (descs, xa_arena, pcm_bytes, payloads)."""
    descs = make_descs(len(specs))
    chunks, payloads = [], []
    xa_off, pcm_off = 0, 0
    for i, s in enumerate(specs):
        blocks = (s["samples"] + 31) // 32
        if "payload" in s:                    # a hand-made stream (extremes())
            pay = np.array(s["payload"], dtype=np.uint8)
            assert pay.size == blocks * s["channels"] * synth.block_size(s["bits"])
        else:
            pay = synth.xa_payload(seed, s.get("key", i), s["bits"], s["channels"],
                                   blocks, s["mix"])
        if "patch" in s:                      # {block-channel index: profile byte}
            bs = synth.block_size(s["bits"])
            for q, val in s["patch"].items():
                pay[q * bs] = val
        gap = (xa_gap * (i + 1)) % 23 if xa_gap else 0
        xa_off += gap
        chunks.append(np.zeros(gap, dtype=np.uint8))
        chunks.append(pay)
        d = descs[i]
        d["xa_off"] = xa_off
        d["pcm_off"] = pcm_off
        d["blocks"] = blocks
        d["pcm_len"] = s["samples"] * 2 * s["channels"]
        d["prev"] = s.get("prev") or ((0, 0), (0, 0))
        d["bits"] = s["bits"]
        d["channels"] = s["channels"]
        xa_off += pay.size
        pcm_off = align16(pcm_off + blocks * 64 * s["channels"]) + 16 * pcm_pad
        payloads.append(pay)
    arena = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return descs, arena, pcm_off, payloads


def pack_codes(bits, codes):
    """32 signed codes of `bits` bits -> the block's payload bytes, MSB first
    (/root/reference/src/libbjxa.c:286-345 read them back the same way)."""
    acc = 0
    for c in codes:
        acc = acc << bits | (int(c) & ((1 << bits) - 1))
    return list(acc.to_bytes(4 * bits, "big"))


def extremes(bits, channels, blocks=40):
    """Streams that drive the predictor to both rails and through every sign of
    the truncating division: full-scale codes (all highest, all lowest, alternating,
    a ramp) at ranges 0..3, every filter 1..4 in chains with no cut block, entry
    states at the rails.  One spec per (filter, pattern, entry state)."""
    hi, lo = (1 << (bits - 1)) - 1, -(1 << (bits - 1))
    pats = {"hi": [hi] * 32, "lo": [lo] * 32, "alt": [hi, lo] * 16, "tla": [lo, hi] * 16,
            "ramp": [lo + (i * (hi - lo)) // 31 for i in range(32)], "ones": [1, -1] * 16}
    states = [((32767, 32767), (-32768, -32768)), ((-32768, 32767), (32767, -32768)),
              ((-1, 0), (0, -1))]
    specs = []
    for f in (1, 2, 3, 4):
        for name, codes in pats.items():
            for k, st in enumerate(states):
                pay = []
                for b in range(blocks):
                    for c in range(channels):
                        # the other channel runs the mirrored pattern, one filter on
                        cc = codes if c == 0 else [-1 - x for x in codes]
                        ff = f if c == 0 else 1 + f % 4
                        pay += [ff << 4 | (b + c + k) % 4] + pack_codes(bits, cc)
                specs.append(dict(bits=bits, channels=channels, samples=32 * blocks - (f + k) % 7,
                                  payload=pay, prev=st, mix="hand-made"))
    return specs


def check_decode(oracle, specs, descs, payloads, pcm_arena, prev_out, first_bad):
    """Compares against the oracle; returns the number of samples checked."""
    total = 0
    for i, s in enumerate(specs):
        d = descs[i]
        blocks = int(d["blocks"])
        done, bad, pcm, st = oracle.decode_blocks(
            s["bits"], s["channels"], np.array(d["prev"]), payloads[i], blocks,
            int(d["pcm_len"]))
        off = int(d["pcm_off"])
        got = pcm_arena[off:off + pcm.size * 2].view(np.int16)
        assert np.array_equal(got, pcm), (i, s, "pcm mismatch at",
                                          int(np.argmax(got != pcm)))
        if bad:
            # first bad block-channel: the oracle stopped inside block `done`
            assert first_bad[i] // s["channels"] == done, (i, s)
        else:
            assert first_bad[i] == 0xFFFFFFFF, (i, s)
            if blocks:
                assert np.array_equal(np.array(prev_out[i])[:s["channels"]],
                                      st[:s["channels"]]), (i, s, prev_out[i], st)
            # nothing written past the PCM owed (guard pattern intact)
            end = off + int(d["pcm_len"])
            nxt = align16(off + blocks * 64 * s["channels"])
            assert (pcm_arena[end:nxt] == 0xCD).all(), (i, s, "overrun")
        total += pcm.size
    return total


def encode_batch(specs, seed=2, xa_gap=0):
    """specs: list of dict(bits, channels, frames, key).  Returns
    (descs, pcm_arena(uint8), xa_bytes, pcms)."""
    descs = make_descs(len(specs))
    chunks, pcms = [], []
    xa_off, pcm_off = 0, 0
    for i, s in enumerate(specs):
        pcm = synth.make_pcm(seed, s.get("key", i), s["channels"], s["frames"])
        blocks = (s["frames"] + 31) // 32
        gap = (xa_gap * (i + 1)) % 23 if xa_gap else 0
        xa_off += gap
        d = descs[i]
        d["xa_off"] = xa_off
        d["pcm_off"] = pcm_off
        d["blocks"] = blocks
        d["pcm_len"] = pcm.size * 2
        d["bits"] = s["bits"]
        d["channels"] = s["channels"]
        raw = pcm.view(np.uint8)
        pad = align16(raw.size) - raw.size
        chunks.append(raw)
        # poison the padding: the kernel must zero-pad, not read neighbours
        chunks.append(np.full(pad, 0x77, dtype=np.uint8))
        pcm_off += raw.size + pad
        xa_off += blocks * s["channels"] * synth.block_size(s["bits"])
        pcms.append(pcm)
    arena = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return descs, arena, xa_off, pcms


def check_encode(oracle, specs, descs, pcms, xa_arena):
    prev_end = 0
    for i, s in enumerate(specs):
        want = oracle.encode_blocks(s["bits"], s["channels"], pcms[i])
        off = int(descs[i]["xa_off"])
        got = xa_arena[off:off + want.size]
        assert np.array_equal(got, want), (i, s, int(np.argmax(got != want)))
        assert (xa_arena[prev_end:off] == 0xCD).all(), (i, s, "wrote into the gap")
        prev_end = off + want.size
    assert (xa_arena[prev_end:] == 0xCD).all()
