#!/usr/bin/env python
"""Single-call latency of bjxa_decode / bjxa_encode through the reference's own API
(include/bjxa.h): microseconds per call for calls of 1 block (the reference CLI's
default mode, /root/reference/src/bjxa_decode.c:102-161), 32 blocks and 1 s of audio,
next to the unmodified reference library on one host core (oracle/_ref, if built).

    python tools/latency_probe.py > profiles/latency_rNN.json

The ctypes call overhead of this probe (about 1-2 us) is inside every figure, on
both sides.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bjxa_b200  # noqa: E402
from bjxa_b200 import synth  # noqa: E402
from oracle import binding  # noqa: E402


def probe(lib, bits, ch, blocks_per_call, calls, mix="P1"):
    total = blocks_per_call * calls
    xa = synth.make_xa(7, 100 + bits + ch, bits, ch, 32 * total, mix=mix)
    dec = lib.decoder()
    assert lib.parse_header(dec, xa[:32]) == 32
    _, fmt = lib.decode_format(dec)
    step = blocks_per_call * fmt.block_size_xa
    buf = np.zeros(blocks_per_call * fmt.block_size_pcm, dtype=np.uint8)
    pay = np.frombuffer(xa, dtype=np.uint8)[32:]
    chunks = [pay[i * step:(i + 1) * step] for i in range(calls)]
    # first call: the thread's staging, the module's kernels
    t0 = time.perf_counter()
    assert lib.decode(dec, buf, buf.size, chunks[0], step) == blocks_per_call
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for c in chunks[1:]:
        lib.decode(dec, buf, buf.size, c, step)
    dt = (time.perf_counter() - t0) / (calls - 1)
    lib.free_decoder(dec)
    return {"us_per_call": round(dt * 1e6, 2), "first_call_us": round(first * 1e6, 1),
            "Msamples_per_s": round(blocks_per_call * 32 * ch / dt / 1e6, 2)}


def main():
    lib = bjxa_b200.load()
    ref = binding.reference_lib() if binding.have_ref() else None
    out = {"what": "bjxa_decode, host buffers, one call after the other on one thread",
           "cases": {}}
    for bits, ch in ((8, 1), (4, 2)):
        for name, bpc, calls in (("1 block", 1, 3000), ("32 blocks", 32, 1000),
                                 ("1 s of audio (1379 blocks)", 1379, 60)):
            key = f"{bits}-bit {'stereo' if ch == 2 else 'mono'}, {name} per call"
            row = {"b200": probe(lib, bits, ch, bpc, calls)}
            if ref is not None:
                row["reference_cpu"] = probe(ref, bits, ch, bpc, calls)
            out["cases"][key] = row
    print(json.dumps(out))


if __name__ == "__main__":
    main()
