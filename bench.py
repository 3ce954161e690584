#!/usr/bin/env python
"""Benchmark of the libbjxa block transform on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[1], SURVEY.md section 8d "Config 2"):
4096 synthetic mono 8-bit XA streams x 60 s @ 44.1 kHz per GPU, generated on the
device.  A "step" is one pass of the decode hot path over the whole batch = one
bjxa_plan_run().  With N > 1 (torchrun, one process per GPU) every rank decodes
its own 4096-stream shard -- the streams are independent, so there is no
data-path collective (weak scaling); only the timing is reduced (max over ranks).

Besides the headline the line carries (unless --no-extras):
  by_profile_mix   the same batch with the other profile-byte mixes
  encode           reference-exact encode of the same batch
  e2e              bjxa_batch_decode on pinned HOST buffers, copies in the timed
                   region, with the copy ceiling of the box measured beside it
  configs          BASELINE.json configs[2] (mixed stereo batch) and configs[3]
                   (stereo 4-bit encode) per GPU, and configs[4]: ONE corpus of
                   262 144 streams sharded by bytes (bjxa_shard_range) over the N
                   ranks -- strong scaling -- decode then encode
  parity           every leg: a checksum of every stream computed on the device
                   (bjxa_plan_checksum), and the oracle on a seeded subset of at
                   least 256 streams whose checksums must agree (all streams with
                   --full-parity); any mismatch fails the run
  cpu_baseline     the unmodified reference on this box's host cores (rank 0)

One JSON line is printed by rank 0; its keys are described in DESIGN.md.
`--impl reference` times the UNMODIFIED reference library (compiled to
oracle/_ref/libbjxa_ref.so) on the host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# ---- workload constants ------------------------------------------------------
N_STREAMS = int(os.environ.get("BJXA_BENCH_STREAMS", 4096))
SECONDS = 60
RATE = 44100
SAMPLES = SECONDS * RATE                       # 2 646 000 per stream
BLOCKS = (SAMPLES + 31) // 32                  # 82 688, last one holds 16 samples
BITS, CH = 8, 1
BS = 4 * BITS + 1
XA_BYTES = BLOCKS * BS * CH                    # 2 728 704
PCM_BYTES = SAMPLES * 2 * CH                   # 5 292 000
PCM_PITCH = (BLOCKS * 64 * CH + 15) & ~15      # 5 292 032
ALGO_BYTES_PER_STREAM = XA_BYTES + PCM_BYTES   # data_len + samples*ch*2
HEADLINE_MIX = "P1"
MIXES = ("P0", "P1", "P2", "P3")
# e2e: the whole batch at N = 1; a shard of it per rank when N ranks share one
# host (8 x 33 GB of pinned memory is more than a bench should take from it)
E2E_STREAMS = int(os.environ.get("BJXA_BENCH_E2E_STREAMS", 0))
CORPUS_STREAMS = int(os.environ.get("BJXA_BENCH_CORPUS_STREAMS", 262144))
PARITY_STREAMS = 256


def mix_profiles(torch, mix, n, blocks, device, seed):
    """(n, blocks) uint8 profile bytes on the device; same distributions as
    bjxa_b200/synth.py:profile_bytes."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = torch.rand((n, blocks), device=device, generator=g)
    v = torch.rand((n, blocks), device=device, generator=g)
    if mix == "P0":
        filt = torch.zeros_like(u, dtype=torch.int32)
        rng = (v * 9).to(torch.int32)
    elif mix == "P1":
        filt = torch.where(u < 0.0066, 1, torch.where(u < 0.0511, 2,
                           torch.where(u < 0.05115, 3, 0))).to(torch.int32)
        for _ in range(4):                    # non-zero filters are isolated
            clash = (filt[:, 1:] != 0) & (filt[:, :-1] != 0)
            if not bool(clash.any()):
                break
            filt[:, 1:][clash] = 0
        hist = torch.tensor([8790, 6891, 3447, 487, 137, 920], device=device,
                            dtype=torch.float32)
        vals = torch.tensor([0, 1, 2, 3, 4, 6], device=device, dtype=torch.int32)
        cum = torch.cumsum(hist, 0) / hist.sum()
        rng = vals[torch.bucketize(v, cum).clamp_(max=5)]
    elif mix == "P2":
        filt = (u * 5).to(torch.int32).clamp_(max=4)
        rng = (v * 16).to(torch.int32).clamp_(max=15)
    elif mix == "P3":
        filt = 1 + (u * 4).to(torch.int32).clamp_(max=3)
        rng = (v * 16).to(torch.int32).clamp_(max=15)
    elif mix.startswith("C"):
        # "Cnn": every block independently a chain block (filter 1..4) with
        # probability nn %, ranges uniform 0..15 -- used to find where one tile
        # form overtakes another
        frac = float(mix[1:]) / 100.0
        w = torch.rand((n, blocks), device=device, generator=g)
        filt = torch.where(u < frac, 1 + (w * 4).to(torch.int32).clamp_(max=3), 0).to(torch.int32)
        rng = (v * 16).to(torch.int32).clamp_(max=15)
    else:
        raise ValueError(mix)
    return ((filt << 4) | rng).to(torch.uint8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region
    (/opt/skills/guides/B200_PROFILING.md, 'clocks' recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes of the decode kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "decode_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


# ---- CPU baseline (the unmodified reference, oracle/_ref) ---------------------

def cpu_reference_rate(xa_streams, steps=1, warmup=0):
    """Decodes `xa_streams` (list of payload arrays, all mono 8-bit 60 s) with
    the reference library on every host core, every stream into a buffer of its
    own; returns (Msamples/s, seconds per pass, info)."""
    from concurrent.futures import ThreadPoolExecutor

    from bjxa_b200 import synth
    from oracle import binding
    if binding.have_ref():
        lib, kind = binding.reference_lib(), "reference"
    else:
        lib, kind = None, "port"
        orc = binding.Oracle()
    cores = os.cpu_count() or 1
    hdr = synth.xa_header(XA_BYTES, SAMPLES, RATE, BITS, CH)
    outs = [np.empty(PCM_PITCH, dtype=np.uint8) for _ in range(len(xa_streams))]

    def job(i):
        pay = xa_streams[i]
        if lib is not None:
            dec = lib.decoder()
            lib.parse_header(dec, hdr)
            got = lib.decode(dec, outs[i], PCM_PITCH, pay, pay.size)   # GIL released
            lib.free_decoder(dec)
            assert got == BLOCKS
        else:
            orc.decode_blocks(BITS, CH, [[0, 0], [0, 0]], pay, BLOCKS, PCM_BYTES)

    def one_pass():
        # static partition: thread t owns streams t, t+cores, ...
        def worker(t):
            for i in range(t, len(xa_streams), cores):
                job(i)
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(worker, range(cores)))

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = (time.perf_counter() - t0) / steps
    rate = len(xa_streams) * SAMPLES * CH / dt / 1e6
    return rate, dt, {"kind": kind, "cores": cores}


def host_sample_streams(n, mix, seed=0xB7A, distinct=32):
    """n streams of the workload on the host: `distinct` generated ones, the rest
    block-rotations of them (same blocks, another order: still independent work)."""
    from bjxa_b200 import synth
    base = [synth.xa_payload(seed, 100000 + i, BITS, CH, BLOCKS, mix)
            for i in range(min(n, distinct))]
    out = []
    for i in range(n):
        b = base[i % len(base)]
        r = (i // len(base)) * 977 % BLOCKS
        out.append(b if r == 0 else np.roll(b.reshape(BLOCKS, BS), r, axis=0).reshape(-1).copy())
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: ~0.02 s of CPU per stream and core -> a fraction of a second per step
    n = max(cores * 32, 64)
    streams = host_sample_streams(n, HEADLINE_MIX)
    rate, dt, info = cpu_reference_rate(streams, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "decode throughput", "value": round(rate, 2),
        "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": workload_config(N_STREAMS),
        "cpu_baseline": {"value": round(rate, 2), "unit": "Msamples/s", "cores": info["cores"],
                         "kind": info["kind"],
                         "sample": f"{n} of the workload's streams per step (32 generated, the "
                                   f"rest block-rotations of them), one reference decoder and "
                                   f"one output buffer per stream, static partition over "
                                   f"{info['cores']} threads"},
        "e2e": {"value": round(rate, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_streams):
    return {"workload": f"batched decode: {n_streams} synthetic mono 8-bit XA streams x "
                        f"{SECONDS} s @ {RATE} Hz per GPU (BASELINE.json configs[1])",
            "streams_per_gpu": n_streams, "samples_per_stream": SAMPLES,
            "blocks_per_stream": BLOCKS, "bits": BITS, "channels": CH,
            "profile_mix": f"{HEADLINE_MIX} (xa.exe-like: 94.9 % filter 0, isolated filters 1-3)",
            "l2": "inputs (11.2 GB) and outputs (21.7 GB) exceed the 126 MB L2; no flush needed",
            "parallelism": "streams sharded by index, one process per GPU, no collective"}


# ---- variable-shape batches (configs[2], configs[4]) ----------------------------

class VarTable:
    """A table of n streams of mixed shapes and lengths -- the same on every rank
    (numpy, seeded): per stream bits, channels, samples, entry state, and the
    algorithmic bytes bjxa_shard_range balances by."""

    def __init__(self, seed, n, bits_choices, ch_choices, sec_lo, sec_hi):
        rng = np.random.default_rng(seed)
        self.n = n
        self.bits = rng.choice(bits_choices, n).astype(np.int64)
        self.ch = rng.choice(ch_choices, n).astype(np.int64)
        secs = np.exp(rng.uniform(np.log(sec_lo), np.log(sec_hi), n))
        self.samples = np.maximum(1, (secs * RATE).astype(np.int64))
        self.blocks = (self.samples + 31) // 32
        self.prev = rng.integers(-32768, 32768, (n, 2, 2))
        self.xa_bytes = self.blocks * self.ch * (4 * self.bits + 1)
        self.pcm_bytes = self.samples * 2 * self.ch
        self.algo = (self.xa_bytes + self.pcm_bytes).astype(np.uint64)


def build_var_batch(torch, dev, tab, lo, hi, mix, seed):
    """Streams [lo, hi) of the table on the device.  The XA arena holds one region
    per bits class (every block of a class has the same size, so the class's
    profile bytes are column 0 of a (blocks, block size) matrix); a stream's
    xa_off points into its class's region.  -> (descs, xa, pcm bytes, samples,
    algorithmic bytes)"""
    from bjxa_b200.api import make_descs
    n = hi - lo
    bits, ch = tab.bits[lo:hi], tab.ch[lo:hi]
    blocks, xa_bytes = tab.blocks[lo:hi], tab.xa_bytes[lo:hi]
    pitch = (blocks * 64 * ch + 15) & ~15
    d = make_descs(n)
    d["pcm_off"] = np.concatenate(([0], np.cumsum(pitch)[:-1]))
    d["blocks"], d["pcm_len"] = blocks, tab.pcm_bytes[lo:hi]
    d["bits"], d["channels"] = bits, ch
    d["prev"] = tab.prev[lo:hi]
    xa_total = int(xa_bytes.sum())
    xa = torch.empty(xa_total + 64, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for o in range(0, xa.numel(), 1 << 30):
        xa[o:o + (1 << 30)].random_(0, 256, generator=g)
    base = 0
    xa_off = np.zeros(n, dtype=np.uint64)
    for b in (4, 6, 8):
        sel = np.nonzero(bits == b)[0]
        if sel.size == 0:
            continue
        bs = 4 * b + 1
        sizes = xa_bytes[sel]
        xa_off[sel] = base + np.concatenate(([0], np.cumsum(sizes)[:-1]))
        total = int(sizes.sum()) // bs            # block-channels of the class
        region = xa[base:base + total * bs].view(total, bs)
        for c0 in range(0, total, 1 << 25):
            c1 = min(total, c0 + (1 << 25))
            region[c0:c1, 0] = mix_profiles(torch, mix, 1, c1 - c0, dev, seed + 31 * b + c0).reshape(-1)
        base += total * bs
    d["xa_off"] = xa_off
    torch.cuda.synchronize()
    return (d, xa, int(pitch.sum()), int((tab.samples[lo:hi] * ch).sum()),
            int(tab.algo[lo:hi].sum()))


# ---- parity inside the timed run ---------------------------------------------

def oracle_pool(jobs):
    """Runs the oracle jobs on all host cores (its C code releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        return list(ex.map(lambda j: j(), jobs))


def parity_decode(lib, plan, d, xa, picks, what):
    """Checksums of every stream's PCM on the device; the oracle decodes `picks`
    and its checksums (and final states) must agree."""
    from bjxa_b200 import synth
    from oracle import binding
    orc = binding.Oracle()
    n = d.size
    res = lib.plan_fetch(plan, n)
    if not (res["result"] == d["blocks"]).all():
        raise SystemExit(f"bench.py: {what}: decode reported failures")
    sums = lib.plan_checksum(plan, n)

    def job(i):
        s = d[i]
        nb = int(s["blocks"]) * int(s["channels"]) * (4 * int(s["bits"]) + 1)
        pay = xa[int(s["xa_off"]):int(s["xa_off"]) + nb].cpu().numpy()
        return lambda: orc.decode_blocks(int(s["bits"]), int(s["channels"]), np.array(s["prev"]),
                                         pay, int(s["blocks"]), int(s["pcm_len"]))
    outs = oracle_pool([job(int(i)) for i in picks])
    for i, (done, bad, want, st) in zip(picks, outs):
        c = int(d[i]["channels"])
        if bad or synth.stream_checksum(want) != int(sums[i]) or \
                not np.array_equal(np.array(res[i]["prev"])[:c], st[:c]):
            raise SystemExit(f"bench.py: {what}: stream {int(i)} differs from the oracle")
    return {"streams_checksummed_on_gpu": int(n),
            "digest_of_all": f"{int(np.bitwise_xor.reduce(sums)):016x}",
            "streams_vs_oracle": int(len(picks)),
            "samples_vs_oracle": int(sum(o[2].size for o in outs)), "ok": True}


def parity_encode(lib, plan, d, pcm, picks, what):
    from bjxa_b200 import synth
    from oracle import binding
    orc = binding.Oracle()
    n = d.size
    lib.plan_fetch(plan, n)
    sums = lib.plan_checksum(plan, n)

    def job(i):
        s = d[i]
        raw = pcm[int(s["pcm_off"]):int(s["pcm_off"]) + int(s["pcm_len"])].cpu().numpy().view(np.int16)
        return lambda: orc.encode_blocks(int(s["bits"]), int(s["channels"]), raw)
    outs = oracle_pool([job(int(i)) for i in picks])
    for i, want in zip(picks, outs):
        if synth.stream_checksum(want) != int(sums[i]):
            raise SystemExit(f"bench.py: {what}: encoded stream {int(i)} differs from the oracle")
    return {"streams_checksummed_on_gpu": int(n),
            "digest_of_all": f"{int(np.bitwise_xor.reduce(sums)):016x}",
            "streams_vs_oracle": int(len(picks)), "ok": True}


def pick_streams(n, full, seed=0xB7A):
    if full or n <= PARITY_STREAMS:
        return np.arange(n)
    return np.sort(np.random.default_rng(seed).choice(n, PARITY_STREAMS, replace=False))


# ---- GPU arm -------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true",
                    help="headline only: skip the per-mix, encode, e2e, configs and CPU legs")
    ap.add_argument("--full-parity", action="store_true",
                    help="the oracle decodes every stream of every leg, not a subset of 256")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import bjxa_b200
    from bjxa_b200.api import PLAN_DECODE, PLAN_ENCODE, make_descs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    out_fd = 1
    if world > 1:
        # stdout carries the one JSON line and nothing else: whatever libraries
        # print on file descriptor 1 ("NCCL version ...") goes to stderr, and
        # rank 0 writes its line to the descriptor saved here
        sys.stdout.flush()
        out_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    lib = bjxa_b200.load()
    warm = max(args.warmup, 3)
    S = N_STREAMS
    peak, peak_src = measured_peak()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def gather(x):
        if world == 1:
            return [float(x)]
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def timed(fn, k):
        """k steps bracketed by barrier + synchronize; device time, max over ranks."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        barrier()
        t0 = time.time()
        ev[0].record()
        for i in range(k):
            fn()
            ev[i + 1].record()
        barrier()
        t1 = time.time()
        own_ms = ev[0].elapsed_time(ev[k])
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(k)]
        return max_over_ranks(own_ms), per, (t0, t1), own_ms

    # every rank owns S streams: global stream index = rank * S + i (weak scaling);
    # bjxa_shard_range gives the same partition for a global table of world*S streams
    first, count = lib.shard_range(world * S, rank, world)
    assert (first, count) == (rank * S, S)

    # ---- synthetic batch, generated on the device ------------------------------
    xa = torch.empty((S, BLOCKS, BS), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(0xB7A + rank)
    for s0 in range(0, S, 128):
        xa[s0:s0 + 128].random_(0, 256, generator=g)

    def set_mix(mix):
        for s0 in range(0, S, 512):
            n = min(512, S - s0)
            xa[s0:s0 + n, :, 0] = mix_profiles(torch, mix, n, BLOCKS, dev,
                                               1000 * rank + s0 + MIXES.index(mix))
        torch.cuda.synchronize()

    pcm = torch.empty(S * PCM_PITCH, dtype=torch.uint8, device=dev)
    descs = make_descs(S)
    descs["xa_off"] = np.arange(S, dtype=np.uint64) * XA_BYTES
    descs["pcm_off"] = np.arange(S, dtype=np.uint64) * PCM_PITCH
    descs["blocks"] = BLOCKS
    descs["pcm_len"] = PCM_BYTES
    descs["bits"], descs["channels"] = BITS, CH
    plan = lib.plan_create(PLAN_DECODE, descs)

    def step():
        lib.plan_run(plan, pcm.data_ptr(), pcm.numel(), xa.data_ptr(), xa.numel(), stream)

    sampler = ClockSampler(local) if rank == 0 else None
    set_mix(HEADLINE_MIX)
    for _ in range(warm):
        step()
    launched0 = lib.plan_launched(plan)
    total_ms, per_launch, window, _ = timed(step, args.steps)
    gpu_launches = lib.plan_launched(plan) - launched0     # counted by the library, launch by launch
    ms_per_step = total_ms / args.steps
    samples_per_step = world * S * SAMPLES * CH
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6
    clocks = sampler.window(*window) if sampler else None

    # ---- parity of what was just timed: every stream checksummed on the device,
    # a seeded subset of them against the oracle (every rank checks its own) ----
    picks = pick_streams(S, args.full_parity, 0xB7A + rank)
    parity = {"headline": parity_decode(lib, plan, descs, xa.reshape(-1), picks, "headline")}

    # ---- roofline of the decode kernel ------------------------------------------
    algo_bytes = S * ALGO_BYTES_PER_STREAM          # per launch, per GPU
    avg_launch_ms = float(np.mean(per_launch))
    achieved = algo_bytes / (avg_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic(),
                "kernel": "xa_decode_kernel<DecTile<8,512,1,3>> (mono, long strips, direct form)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                "avg_launch_ms": round(avg_launch_ms, 4),
                "frac_of_nominal_8000": round(achieved / 8000.0, 4)}

    line = {
        "metric": "decode throughput", "value": round(value, 1), "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(S), "roofline": roofline,
        "gpu_launches": gpu_launches, "clocks": clocks, "parity": parity,
    }

    if not args.no_extras:
        # ---- the other profile mixes (each reported separately) ----
        by_mix = {HEADLINE_MIX: {"Msamples_per_s": round(value, 1),
                                 "hbm_frac": roofline["frac"]}}
        for mix in MIXES:
            if mix == HEADLINE_MIX:
                continue
            set_mix(mix)
            step()
            step()
            k = 5
            tms, _, _, _ = timed(step, k)
            v = samples_per_step / (tms / k * 1e-3) / 1e6
            by_mix[mix] = {"Msamples_per_s": round(v, 1),
                           "hbm_frac": round(v * 1e6 / world * ALGO_BYTES_PER_STREAM / SAMPLES
                                             / 1e9 / peak, 4)}
            if mix == "P2":     # the chain-rich mix: also checked against the oracle
                parity["P2"] = parity_decode(lib, plan, descs, xa.reshape(-1), picks[:64], "mix P2")
        line["by_profile_mix"] = by_mix
        set_mix(HEADLINE_MIX)
        step()
        torch.cuda.synchronize()

        # ---- encode (reference-exact) over the PCM just produced ------------------
        eplan = lib.plan_create(PLAN_ENCODE, descs)
        xa_out = torch.empty(S * XA_BYTES + 16, dtype=torch.uint8, device=dev)

        def estep():
            lib.plan_run(eplan, xa_out.data_ptr(), xa_out.numel(), pcm.data_ptr(),
                         pcm.numel(), stream)
        for _ in range(3):
            estep()
        tms, _, _, _ = timed(estep, 5)
        ems = tms / 5
        parity["encode"] = parity_encode(lib, eplan, descs, pcm, picks, "encode")
        line["encode"] = {"Msamples_per_s": round(samples_per_step / (ems * 1e-3) / 1e6, 1),
                          "ms_per_step": round(ems, 4), "mode": "reference-exact (profile 0)",
                          "hbm_frac": round(algo_bytes / (ems * 1e-3) / 1e9 / peak, 4)}
        lib.plan_free(eplan)
        del xa_out

        # ---- end to end through the host-buffer C-ABI call ------------------------
        ne = E2E_STREAMS or (S if world == 1 else 512)
        ne = min(ne, S)
        try:
            h_src = torch.empty((ne, XA_BYTES), dtype=torch.uint8, pin_memory=True)
            h_dst = torch.empty((ne, PCM_PITCH), dtype=torch.uint8, pin_memory=True)
        except RuntimeError:            # the host cannot pin 33 GB: a shard of the batch
            ne = min(512, S)
            h_src = torch.empty((ne, XA_BYTES), dtype=torch.uint8, pin_memory=True)
            h_dst = torch.empty((ne, PCM_PITCH), dtype=torch.uint8, pin_memory=True)
        for s0 in range(0, ne, 256):
            h_src[s0:s0 + 256].copy_(xa[s0:s0 + 256].reshape(-1, XA_BYTES))
        from bjxa_b200 import synth
        hdr = synth.xa_header(XA_BYTES, SAMPLES, RATE, BITS, CH)
        srcs = [h_src[i].numpy() for i in range(ne)]
        dsts = [h_dst[i].numpy() for i in range(ne)]

        def e2e_step():
            decs = []
            for _ in range(ne):
                d = lib.decoder()
                lib.parse_header(d, hdr)
                decs.append(d)
            res, errs = lib.batch_decode(decs, dsts, srcs)
            for d in decs:
                lib.free_decoder(d)
            assert all(r == BLOCKS for r in res)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        k = 3 if ne <= 512 else 2
        for _ in range(k):
            e2e_step()
        barrier()
        edt = max_over_ranks((time.perf_counter() - t0) / k)
        for i in (0, ne // 2, ne - 1):
            want = pcm[i * PCM_PITCH:i * PCM_PITCH + PCM_BYTES].cpu().numpy()
            assert np.array_equal(h_dst[i].numpy()[:PCM_BYTES], want), "e2e output differs"
        # the box's copy ceiling for the same bytes: both directions at once on two
        # streams, all ranks together, nothing else
        d_probe_in = xa.reshape(-1)[:ne * XA_BYTES]
        d_probe_out = pcm[:ne * PCM_PITCH]
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def copies():
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                d_probe_in.copy_(h_src.reshape(-1), non_blocking=True)
            with torch.cuda.stream(s2):
                h_dst.reshape(-1).copy_(d_probe_out, non_blocking=True)
            cur.wait_stream(s1)
            cur.wait_stream(s2)
        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            copies()
        barrier()
        cdt = max_over_ranks((time.perf_counter() - t0) / 2)
        e2e_val = world * ne * SAMPLES * CH / edt / 1e6
        ceiling = world * ne * SAMPLES * CH / cdt / 1e6
        line["e2e"] = {"value": round(e2e_val, 1),
                       "unit": "Msamples/s", "h2d_bytes_per_step": ne * XA_BYTES,
                       "d2h_bytes_per_step": ne * PCM_BYTES,
                       "copy_ceiling_Msamples_per_s": round(ceiling, 1),
                       "frac_of_box_ceiling": round(e2e_val / ceiling, 4),
                       "copy_ceiling_GBps_per_gpu": round(ne * (XA_BYTES + PCM_PITCH) / cdt / 1e9, 2),
                       "call": f"bjxa_batch_decode on {ne} streams per GPU per step, pinned "
                               f"host buffers, copies inside the timed region; ceiling = the "
                               f"same bytes copied both ways at once by all {world} rank(s), "
                               f"nothing else running"}
        del h_src, h_dst, srcs, dsts, d_probe_in, d_probe_out

    # the headline's arenas are no longer needed
    lib.plan_free(plan)
    del xa, pcm
    torch.cuda.empty_cache()

    if not args.no_extras:
        configs = {}

        # ---- configs[2]: mixed stereo batch, per GPU (weak) ------------------------
        tab = VarTable(3 + rank, 4096, [4, 6, 8], [2], 0.5, 120.0)
        d, vxa, pcm_total, nsamp, algo = build_var_batch(torch, dev, tab, 0, tab.n, "P2", 300 + rank)
        vpcm = torch.empty(pcm_total + 64, dtype=torch.uint8, device=dev)
        vplan = lib.plan_create(PLAN_DECODE, d)

        def vstep():
            lib.plan_run(vplan, vpcm.data_ptr(), vpcm.numel(), vxa.data_ptr(), vxa.numel(), stream)
        vstep()
        vstep()
        vl0 = lib.plan_launched(vplan)
        tms, _, _, _ = timed(vstep, 3)
        vlaunches = (lib.plan_launched(vplan) - vl0) / 3
        ms = tms / 3
        tot_samples = sum_over_ranks(nsamp)
        par = parity_decode(lib, vplan, d, vxa, pick_streams(tab.n, args.full_parity, 7), "configs[2]")
        configs["configs[2]"] = {
            "what": "mixed batch decode: 4096 stereo streams per GPU, 4/6/8 bit, 0.5-120 s "
                    "log-uniform, mix P2, random header state",
            "ms_per_step": round(ms, 3), "launches_per_step": round(vlaunches, 1),
            "Msamples_per_s": round(tot_samples / ms / 1e3, 1),
            "hbm_frac": round(algo / (ms * 1e-3) / 1e9 / peak, 4), "parity": par}
        lib.plan_free(vplan)
        del vxa, vpcm
        torch.cuda.empty_cache()

        # ---- configs[3]: 4096 stereo PCM streams x 60 s -> 4-bit XA, per GPU --------
        n3, ch3, bits3 = 4096, 2, 4
        pcm_bytes3 = SAMPLES * 2 * ch3
        pitch3 = (BLOCKS * 64 * ch3 + 15) & ~15
        xa_bytes3 = BLOCKS * ch3 * (4 * bits3 + 1)
        pcm3 = torch.empty((n3, pitch3 // 2), dtype=torch.int16, device=dev)
        g3 = torch.Generator(device=dev)
        g3.manual_seed(4 + rank)
        t = torch.arange(pitch3 // 4, device=dev, dtype=torch.int32)
        for s0 in range(0, n3, 128):
            m = min(128, n3 - s0)
            # integer-only PCM: a triangle wave + noise per channel
            per = torch.randint(16, 2000, (m, 1), device=dev, generator=g3, dtype=torch.int32)
            amp = torch.randint(3000, 12000, (m, 1), device=dev, generator=g3, dtype=torch.int32)
            ph = t.unsqueeze(0) % per
            tri = ((2 * ph - per).abs() * 2 - per) * amp // per
            noise = torch.randint(-2048, 2048, (m, pitch3 // 4), device=dev, generator=g3,
                                  dtype=torch.int32)
            pcm3[s0:s0 + m, 0::2] = (tri + noise).clamp_(-32768, 32767).to(torch.int16)
            pcm3[s0:s0 + m, 1::2] = (tri // 2 - noise).clamp_(-32768, 32767).to(torch.int16)
            del ph, tri, noise
        d3 = make_descs(n3)
        d3["xa_off"] = np.arange(n3, dtype=np.uint64) * xa_bytes3
        d3["pcm_off"] = np.arange(n3, dtype=np.uint64) * pitch3
        d3["blocks"], d3["pcm_len"] = BLOCKS, pcm_bytes3
        d3["bits"], d3["channels"] = bits3, ch3
        xa3 = torch.empty(n3 * xa_bytes3 + 64, dtype=torch.uint8, device=dev)
        raw3 = pcm3.view(torch.uint8).reshape(-1)
        plan3 = lib.plan_create(PLAN_ENCODE, d3)

        def step3():
            lib.plan_run(plan3, xa3.data_ptr(), xa3.numel(), raw3.data_ptr(), raw3.numel(), stream)
        step3()
        step3()
        tms, _, _, _ = timed(step3, 3)
        ms = tms / 3
        par = parity_encode(lib, plan3, d3, raw3, pick_streams(n3, args.full_parity, 9), "configs[3]")
        configs["configs[3]"] = {
            "what": "batched encode: 4096 stereo PCM streams x 60 s per GPU -> 4-bit XA, "
                    "reference-exact (the reference writes profile 0: libbjxa.c:679; the "
                    "searching encoder is an opt-in extension, tools/bench_configs.py)",
            "ms_per_step": round(ms, 3),
            "Msamples_per_s": round(world * n3 * SAMPLES * ch3 / ms / 1e3, 1),
            "hbm_frac": round(n3 * (pcm_bytes3 + xa_bytes3) / (ms * 1e-3) / 1e9 / peak, 4),
            "parity": par}
        lib.plan_free(plan3)
        del pcm3, raw3, xa3
        torch.cuda.empty_cache()

        # ---- configs[4]: ONE corpus, sharded by bytes over the ranks (strong) --------
        def corpus(tag, tab, mix):
            lo, cnt = lib.shard_range(tab.n, rank, world, tab.algo)
            d, cxa, pcm_total, nsamp, algo = build_var_batch(torch, dev, tab, lo, lo + cnt, mix, 500)
            cpcm = torch.empty(pcm_total + 64, dtype=torch.uint8, device=dev)
            dplan = lib.plan_create(PLAN_DECODE, d)

            def dstep():
                lib.plan_run(dplan, cpcm.data_ptr(), cpcm.numel(), cxa.data_ptr(), cxa.numel(), stream)
            dstep()
            dstep()
            dl0 = lib.plan_launched(dplan)
            tms, _, _, own = timed(dstep, 3)
            launches = round((lib.plan_launched(dplan) - dl0) / 3, 1)
            dms = tms / 3
            per_rank = gather(own / 3)
            par_d = parity_decode(lib, dplan, d, cxa, pick_streams(cnt, args.full_parity, 11 + rank),
                                  f"configs[4] {tag} decode")
            lib.plan_free(dplan)
            # ... then encode the PCM just produced back to XA (same shapes)
            cxa2 = torch.empty(cxa.numel(), dtype=torch.uint8, device=dev)
            eplan = lib.plan_create(PLAN_ENCODE, d)

            def estep():
                lib.plan_run(eplan, cxa2.data_ptr(), cxa2.numel(), cpcm.data_ptr(), cpcm.numel(), stream)
            estep()
            estep()
            tms, _, _, eown = timed(estep, 3)
            ems = tms / 3
            eper_rank = gather(eown / 3)
            par_e = parity_encode(lib, eplan, d, cpcm, pick_streams(cnt, args.full_parity, 13 + rank),
                                  f"configs[4] {tag} encode")
            lib.plan_free(eplan)
            total_samples = int((tab.samples * tab.ch).sum())
            total_algo = float(tab.algo.sum())
            shard_bytes = gather(float(algo))
            out = {
                "streams": int(tab.n), "mix": mix, "samples": total_samples,
                "algorithmic_GB": round(total_algo / 1e9, 2),
                "sharding": "bjxa_shard_range(bytes[]): contiguous, byte-balanced ranges of ONE "
                            "global table, no collective (strong scaling)",
                "streams_this_rank": int(cnt), "launches_per_step": launches,
                "shard_bytes_imbalance": round(max(shard_bytes) / (sum(shard_bytes) / world) - 1, 4),
                "decode": {"ms_per_step": round(dms, 3), "per_rank_ms": [round(x, 3) for x in per_rank],
                           "time_imbalance": round(max(per_rank) / (sum(per_rank) / world) - 1, 4),
                           "Msamples_per_s": round(total_samples / dms / 1e3, 1),
                           "hbm_frac_per_gpu": round(total_algo / world / (dms * 1e-3) / 1e9 / peak, 4),
                           "parity": par_d},
                "encode": {"ms_per_step": round(ems, 3), "per_rank_ms": [round(x, 3) for x in eper_rank],
                           "time_imbalance": round(max(eper_rank) / (sum(eper_rank) / world) - 1, 4),
                           "Msamples_per_s": round(total_samples / ems / 1e3, 1),
                           "hbm_frac_per_gpu": round(total_algo / world / (ems * 1e-3) / 1e9 / peak, 4),
                           "parity": par_e}}
            del cxa, cxa2, cpcm
            torch.cuda.empty_cache()
            return out

        # SURVEY.md 8d "Config 5": the shapes of config 3 (stereo, 4/6/8 bit, mix P2,
        # random header state), 0.25-4 s; beside it the xa.exe-like mix over mono and
        # stereo streams that round 1 reported
        configs["configs[4]"] = {
            "what": f"corpus decode+encode: {CORPUS_STREAMS} streams, 0.25-4 s log-uniform, "
                    f"sharded by bytes over {world} GPU(s)",
            "stereo_P2": corpus("stereo P2", VarTable(5, CORPUS_STREAMS, [4, 6, 8], [2], 0.25, 4.0), "P2"),
            "mixed_P1": corpus("mixed P1", VarTable(6, CORPUS_STREAMS, [4, 6, 8], [1, 2], 0.25, 4.0), "P1")}
        line["configs"] = configs

        # ---- CPU baseline: the reference on this box's host cores (rank 0) ----------
        if rank == 0:
            cores = os.cpu_count() or 1
            # bounded sample: about 10-30 s of CPU work (64 streams = 0.17 Gsamples a core)
            n = 64 * cores
            streams = host_sample_streams(n, HEADLINE_MIX, distinct=16)
            cpu_reference_rate(streams[:cores], steps=1, warmup=0)      # thread start-up, page-in
            rate, dt, info = cpu_reference_rate(streams, steps=1, warmup=0)
            line["cpu_baseline"] = {
                "value": round(rate, 2), "unit": "Msamples/s", "cores": info["cores"],
                "kind": info["kind"],
                "sample": f"{n} streams of this workload (16 generated, the rest block-rotations "
                          f"of them) decoded once, one reference decoder and one output buffer "
                          f"per stream, static partition over {info['cores']} threads, "
                          f"{dt:.2f} s wall = {dt * info['cores']:.0f} core-seconds"}

    if sampler:
        sampler.stop()
    barrier()
    if rank == 0:
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
