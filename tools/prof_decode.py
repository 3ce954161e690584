#!/usr/bin/env python
"""Small timing / profiling driver for one decode (or encode) launch shape.

    python tools/prof_decode.py --mix P1 --streams 512 --steps 3 [--bits 8 --ch 1]

Generates the batch on the device (same generator as bench.py), runs it through
bjxa_plan_run and prints one JSON line per mix with CUDA-event timings.  Used
under `ncu` for the profiles/ captures and for tile-geometry sweeps
(BJXA_LIB=path selects an alternative build of the library).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import bjxa_b200  # noqa: E402
from bjxa_b200.api import PLAN_DECODE, PLAN_ENCODE, PLAN_ENCODE_SEARCH, Bjxa, make_descs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mix", default="P1")
    ap.add_argument("--streams", type=int, default=512)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--ch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--encode", action="store_true")
    ap.add_argument("--search", action="store_true",
                    help="also run the searching encoder (BJXA_PLAN_ENCODE_SEARCH) on the decoded PCM")
    ap.add_argument("--tag", default="")
    ap.add_argument("--padx", type=int, default=0, help="bytes between consecutive streams' XA payloads")
    ap.add_argument("--padp", type=int, default=0, help="extra bytes (x16) between consecutive streams' PCM")
    a = ap.parse_args()

    lib = Bjxa(os.environ["BJXA_LIB"]) if os.environ.get("BJXA_LIB") else bjxa_b200.load()
    dev = torch.device("cuda", 0)
    S, bits, ch = a.streams, a.bits, a.ch
    samples = int(a.seconds * 44100)
    blocks = (samples + 31) // 32
    bs = (4 * bits + 1) * ch
    xa_bytes = blocks * bs
    pcm_bytes = samples * 2 * ch
    pitch = ((blocks * 64 * ch + 15) & ~15) + 16 * a.padp
    xrow = xa_bytes + a.padx
    xa_flat = torch.zeros((S, xrow), dtype=torch.uint8, device=dev)
    xa = xa_flat[:, :xa_bytes].view(S, blocks * ch, 4 * bits + 1)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    for s0 in range(0, S, 128):
        xa[s0:s0 + 128].random_(0, 256, generator=g)
    pcm = torch.empty(S * pitch, dtype=torch.uint8, device=dev)
    descs = make_descs(S)
    descs["xa_off"] = np.arange(S, dtype=np.uint64) * xrow
    descs["pcm_off"] = np.arange(S, dtype=np.uint64) * pitch
    descs["blocks"], descs["pcm_len"] = blocks, pcm_bytes
    descs["bits"], descs["channels"] = bits, ch
    stream = torch.cuda.current_stream().cuda_stream
    algo = S * (xa_bytes + pcm_bytes)
    assert not (a.encode or a.search) or a.padx == 0

    for mix in a.mix.split(","):
        for s0 in range(0, S, 256):
            n = min(256, S - s0)
            # one profile sequence per stream-CHANNEL (filters of the two channels
            # of a stereo stream are independent), interleaved L,R,L,R like the blocks
            pr = bench.mix_profiles(torch, mix, n * ch, blocks, dev, s0 + 7)
            xa[s0:s0 + n, :, 0] = pr.view(n, ch, blocks).transpose(1, 2).reshape(n, blocks * ch)
        plan = lib.plan_create(PLAN_DECODE, descs)

        def step():
            lib.plan_run(plan, pcm.data_ptr(), pcm.numel(), xa_flat.data_ptr(), xa_flat.numel(), stream)
        for _ in range(a.warmup):
            step()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(a.steps):
            step()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
        res = lib.plan_fetch(plan, S)
        assert (res["result"] == blocks).all()
        best = min(ms)
        out = {"tag": a.tag, "kind": "decode", "mix": mix, "bits": bits, "ch": ch, "streams": S,
               "ms": [round(m, 4) for m in ms],
               "Gsamples_s": round(S * samples * ch / best / 1e6, 1),
               "GBps": round(algo / best / 1e6, 1), "frac_6550": round(algo / best / 1e6 / 6550.4, 4)}
        print(json.dumps(out), flush=True)
        lib.plan_free(plan)

    if a.encode:
        eplan = lib.plan_create(PLAN_ENCODE, descs)
        xo = torch.empty(S * xa_bytes + 16, dtype=torch.uint8, device=dev)

        def estep():
            lib.plan_run(eplan, xo.data_ptr(), xo.numel(), pcm.data_ptr(), pcm.numel(), stream)
        estep()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(a.steps):
            estep()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
        best = min(ms)
        print(json.dumps({"tag": a.tag, "kind": "encode", "bits": bits, "ch": ch, "streams": S,
                          "ms": [round(m, 4) for m in ms],
                          "Gsamples_s": round(S * samples * ch / best / 1e6, 1),
                          "GBps": round(algo / best / 1e6, 1),
                          "frac_6550": round(algo / best / 1e6 / 6550.4, 4)}), flush=True)
        lib.plan_free(eplan)

    if a.search:
        splan = lib.plan_create(PLAN_ENCODE_SEARCH, descs)
        xo = torch.empty(S * xa_bytes + 16, dtype=torch.uint8, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        lib.plan_run(splan, xo.data_ptr(), xo.numel(), pcm.data_ptr(), pcm.numel(), stream)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        print(json.dumps({"tag": a.tag, "kind": "search", "bits": bits, "ch": ch, "streams": S,
                          "ms": round(ms, 2), "Gsamples_s": round(S * samples * ch / ms / 1e6, 2),
                          "Gcandidate_samples_s": round(5 * (17 - bits) * S * samples * ch / ms / 1e6, 1)}),
              flush=True)
        lib.plan_free(splan)


if __name__ == "__main__":
    main()
