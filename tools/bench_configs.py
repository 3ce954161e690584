#!/usr/bin/env python
"""What bench.py's line does not carry (BASELINE.json configs[1..4] are all in
bench.py since round 2; configs[0] is tests/test_gpu_decode.py::
test_config1_square_stereo_4):

    python tools/bench_configs.py [--only encode,files] > profiles/extras_rNN.json

  encode  configs[3]'s batch (4096 stereo PCM streams x 60 s -> 4-bit XA) through
          the reference-exact encoder and then through the SEARCHING encoder, an
          opt-in extension: the reference has no filter/range search
          (/root/reference/src/libbjxa.c:679)
  files   40 000 whole .xa files in a pinned host arena -> whole .wav files
          (bjxa_corpus_run), wall clock including the copies

Each leg prints one JSON line: device-resident throughput (CUDA events around
bjxa_plan_run), algorithmic GB/s and fraction of the measured HBM peak, and a
parity spot check of a few streams against the oracle.
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
from bjxa_b200.api import PLAN_DECODE, PLAN_ENCODE, PLAN_ENCODE_SEARCH, make_descs  # noqa: E402
from oracle import binding  # noqa: E402

DEV = torch.device("cuda", 0)
PEAK, PEAK_SRC = bench.measured_peak()


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]


def leg_encode(lib, orc):
    n, ch, bits = 4096, 2, 4
    samples = 60 * 44100
    blocks = (samples + 31) // 32
    pcm_bytes = samples * 2 * ch
    pitch = (blocks * 64 * ch + 15) & ~15
    xa_bytes = blocks * ch * (4 * bits + 1)
    # integer-only PCM on the device: triangle waves + noise
    pcm = torch.empty((n, pitch // 2), dtype=torch.int16, device=DEV)
    g = torch.Generator(device=DEV)
    g.manual_seed(4)
    t = torch.arange(pitch // 4, device=DEV, dtype=torch.int32)
    for s0 in range(0, n, 64):
        m = min(64, n - s0)
        per = torch.randint(16, 2000, (m, 1), device=DEV, generator=g, dtype=torch.int32)
        amp = torch.randint(3000, 12000, (m, 1), device=DEV, generator=g, dtype=torch.int32)
        ph = t.unsqueeze(0) % per
        tri = ((2 * ph - per).abs() * 2 - per) * amp // per
        noise = torch.randint(-2048, 2048, (m, pitch // 4), device=DEV, generator=g,
                              dtype=torch.int32)
        left = (tri + noise).clamp_(-32768, 32767).to(torch.int16)
        right = (tri // 2 - noise).clamp_(-32768, 32767).to(torch.int16)
        pcm[s0:s0 + m, 0::2] = left
        pcm[s0:s0 + m, 1::2] = right
        del ph, tri, noise, left, right
    d = make_descs(n)
    d["xa_off"] = np.arange(n, dtype=np.uint64) * xa_bytes
    d["pcm_off"] = np.arange(n, dtype=np.uint64) * pitch
    d["blocks"], d["pcm_len"] = blocks, pcm_bytes
    d["bits"], d["channels"] = bits, ch
    xa = torch.empty(n * xa_bytes + 64, dtype=torch.uint8, device=DEV)
    plan = lib.plan_create(PLAN_ENCODE, d)
    st = torch.cuda.current_stream().cuda_stream
    raw = pcm.view(torch.uint8).reshape(-1)
    ms = timed(lambda: lib.plan_run(plan, xa.data_ptr(), xa.numel(), raw.data_ptr(),
                                    raw.numel(), st), 3)
    lib.plan_fetch(plan, n)
    for i in (0, n // 3, n - 1):
        want = orc.encode_blocks(bits, ch, pcm[i].cpu().numpy()[:samples * ch])
        got = xa[i * xa_bytes:(i + 1) * xa_bytes].cpu().numpy()
        assert np.array_equal(got, want), f"encoded stream {i} differs from the oracle"
    best = min(ms)
    algo = n * (pcm_bytes + xa_bytes)
    out = {"config": "configs[3] batched encode: 4096 stereo PCM streams x 60 s -> 4-bit XA, "
                     "reference-exact (the reference has no filter/range search)",
           "ms": [round(m, 3) for m in ms],
           "Msamples_per_s": round(n * samples * ch / best / 1e3, 1),
           "GBps": round(algo / best / 1e6, 1), "hbm_frac": round(algo / best / 1e6 / PEAK, 4),
           "bytes": {"pcm": n * pcm_bytes, "xa": n * xa_bytes}, "parity": "3 streams vs oracle ok"}
    lib.plan_free(plan)

    # The same batch through the searching encoder (an extension: filter 0-4 x
    # range 0-12 per block, closed loop; include/bjxa_batch.h).  Compute-bound --
    # 65 simulations per sample -- so one step, and the oracle check on the first
    # seconds of three streams (the CPU restatement manages 0.8 Msamples/s a core).
    plan = lib.plan_create(PLAN_ENCODE_SEARCH, d)
    sms = timed(lambda: lib.plan_run(plan, xa.data_ptr(), xa.numel(), raw.data_ptr(),
                                     raw.numel(), st), 1, warmup=0)
    res = lib.plan_fetch(plan, n)
    assert (res["result"] == d["blocks"]).all()
    head = 32 * 400
    for i in (0, n // 3, n - 1):
        want, _ = orc.encode_search_blocks(bits, ch, [[0, 0], [0, 0]],
                                           pcm[i].cpu().numpy()[:head * ch])
        got = xa[i * xa_bytes:i * xa_bytes + want.size].cpu().numpy()
        assert np.array_equal(got, want), f"searched stream {i} differs from the oracle"
    sbest = min(sms)
    out["search"] = {"what": "same batch, BJXA_PLAN_ENCODE_SEARCH (extension; 65 candidates per block)",
                     "ms": [round(m, 2) for m in sms],
                     "Msamples_per_s": round(n * samples * ch / sbest / 1e3, 1),
                     "Gcandidate_samples_per_s": round(65 * n * samples * ch / sbest / 1e6, 1),
                     "parity": "first 400 blocks of 3 streams vs oracle ok"}
    lib.plan_free(plan)
    return out


def leg_files(lib, orc):
    """Whole .xa files in a pinned host arena -> whole .wav files in another
    (bjxa_corpus_run): header parse, one copy per chunk each way, files
    assembled on the device.  Wall clock of the call, copies included."""
    import ctypes as C
    import time

    from bjxa_b200 import synth
    from bjxa_b200.api import CORPUS_XA_TO_WAV, FILE_DTYPE
    rng = np.random.default_rng(11)
    protos = []
    for i in range(48):
        bits, ch = (4, 6, 8)[i % 3], 1 + (i // 3) % 2
        secs = float(np.exp(rng.uniform(np.log(0.25), np.log(4.0))))
        protos.append(synth.make_xa(0xB7A, 5000 + i, bits, ch, int(secs * 44100), mix="P1"))
    n = 40000
    pick = rng.integers(0, len(protos), n)
    table = np.zeros(n, dtype=FILE_DTYPE)
    lens = np.array([len(protos[k]) for k in pick], dtype=np.uint64)
    table["in_len"] = lens
    table["in_off"] = np.concatenate(([0], np.cumsum(lens)[:-1]))
    in_bytes = int(lens.sum())
    h_in = lib._bjxa_host_alloc(in_bytes)
    arena = np.ctypeslib.as_array((C.c_uint8 * in_bytes).from_address(h_in))
    for k, t in zip(pick, table):
        o = int(t["in_off"])
        arena[o:o + int(t["in_len"])] = np.frombuffer(protos[k], dtype=np.uint8)
    need = C.c_uint64(0)
    assert lib._bjxa_corpus_extent(CORPUS_XA_TO_WAV, h_in, in_bytes, table.ctypes.data, n,
                                   C.byref(need)) == 0
    out_bytes = int(need.value)
    h_out = lib._bjxa_host_alloc(out_bytes)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        rc = lib._bjxa_corpus_run(CORPUS_XA_TO_WAV, h_in, in_bytes, h_out, out_bytes,
                                  table.ctypes.data, n)
        times.append(time.perf_counter() - t0)
        assert rc == 0, f"bjxa_corpus_run: errno {lib.errno()}"
        assert (table["error"] == 0).all(), np.unique(table["error"], return_counts=True)
    out = np.ctypeslib.as_array((C.c_uint8 * out_bytes).from_address(h_out))
    for i in (0, n // 2, n - 1):
        t = table[i]
        got = out[int(t["out_off"]):int(t["out_off"] + t["out_len"])].tobytes()
        assert got == orc.xa_to_wav(protos[pick[i]]), f"file {i} differs from the oracle"
    samples = int(((table["out_len"].astype(np.int64) - 44) // 2).sum())
    best = min(times)
    res = {"config": "files: 40000 whole .xa files (mono/stereo, 4/6/8 bit, 0.25-4 s, mix P1) in a "
                     "pinned arena -> whole .wav files, bjxa_corpus_run, wall clock incl. copies",
           "s": [round(t, 4) for t in times], "in_GB": round(in_bytes / 1e9, 3),
           "out_GB": round(out_bytes / 1e9, 3),
           "Msamples_per_s": round(samples / best / 1e6, 1),
           "pcie_GBps_in_plus_out": round((in_bytes + out_bytes) / best / 1e9, 2),
           "files_per_s": round(n / best), "parity": "3 files vs oracle ok"}
    lib._bjxa_host_free(h_in)
    lib._bjxa_host_free(h_out)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="encode,files")
    a = ap.parse_args()
    lib = bjxa_b200.load()
    orc = binding.Oracle()
    legs = a.only.split(",")
    print(json.dumps({"peak_GBps": PEAK, "peak_source": PEAK_SRC}), flush=True)
    if "encode" in legs:
        print(json.dumps(leg_encode(lib, orc)), flush=True)
        torch.cuda.empty_cache()
    if "files" in legs:
        print(json.dumps(leg_files(lib, orc)), flush=True)


if __name__ == "__main__":
    main()
