#!/usr/bin/env python
"""Turns what tools/refresh_profiles.sh left in gpurun_out/ into the committed
summaries under profiles/ (run here, after the gpurun call has merged its files):

    python tools/summarize_profiles.py r1

  bench_<tag>.json, bench_ref_<tag>.json, configs_<tag>.json, pcie_<tag>.json   copied
  launches_<tag>_summary.md        from launches_<tag>.csv (ncu launch list)
  decode_p1_4096_<tag>_ncu_summary.csv, decode_traffic.json
                                   from decode_p1_4096_<tag>.ncu-rep (ncu --set full)
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def fnum(x):
    return float(x.replace(",", ""))


def copy(name):
    src = os.path.join(OUT, name)
    if os.path.exists(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(PROF, name))
        return True
    print("missing:", name)
    return False


def launches(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        print("missing:", path)
        return
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per, order = OrderedDict(), []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = fnum(r["Metric Value"])
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        k = r["Kernel Name"]
        per.setdefault(k, [0, 0.0])
        per[k][0] += 1
        per[k][1] += ms
        order.append((k, ms))
    total = sum(v[1] for v in per.values())
    cmd = "python bench.py --steps 2 --warmup 3 --no-extras"
    L = [f"# ncu launch list, round {tag} -- `{cmd}`", "",
         f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "
         f"gpurun_out/launches_{tag}.csv {cmd}` on one B200 (tools/refresh_profiles.sh); the same "
         "command had exited 0 without ncu directly before.",
         "Times are ncu's serialised, cold-cache per-launch durations: compare SHARES, not absolutes.",
         "",
         f"{len(order)} launches, {total:.2f} ms in total. Everything that is not `xa_*` is the "
         "synthetic-input generator (torch RNG / indexing kernels that build the 11.2 GB XA arena and "
         "its profile bytes) and runs BEFORE the timed region.", "",
         "| kernel | launches | total ms | share of all launches |", "|---|---:|---:|---:|"]
    for k, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1])[:10]:
        L.append(f"| `{k[:90]}` | {n} | {ms:.3f} | {100 * ms / total:.1f} % |")
    ours = [(k, ms) for k, ms in order if "xa_" in k]
    L += ["", "## The timed region", "",
          "One step = one `bjxa_plan_run()` = one `cudaMemsetAsync` (arms first_bad[], the ticket "
          "counters and the census words; a memset node, not a kernel) + the launches below. The "
          "4096-stream class has several candidate forms (long strips, wide tiles, the relay form's "
          "two passes, the segment form), so the census kernel runs and the forms it does not pick return at once:", "",
          "| launch | kernel | ms under ncu |", "|---|---|---:|"]
    n_run = sum(1 for k, _ in ours if "xa_checksum" not in k)     # the parity checksum follows the timed region
    per_step = n_run // 5 if n_run % 5 == 0 else None
    for i, (k, ms) in enumerate(ours):
        step = i // per_step if per_step else i
        kind = "warm-up" if step < 3 else "timed"
        L.append(f"| {i} ({kind} step {step}) | `{k[:100]}` | {ms:.4f} |")
    if per_step:
        timed = [(k, ms) for k, ms in ours[3 * per_step:] if "xa_checksum" not in k]
        tot = sum(ms for _, ms in timed)
        dec = sum(ms for k, ms in timed if "xa_decode_kernel" in k)
        big = max(ms for _, ms in timed)
        L += ["", f"Share of the dominant `xa_decode_kernel` launch in a timed step: "
              f"{100 * big * 2 / tot:.1f} % of the step's kernel time "
              f"(all `xa_decode_kernel` launches {100 * dec / tot:.1f} %; the census and the candidate forms it did not pick are the rest)."]
    try:
        b = json.loads(open(os.path.join(PROF, f"bench_{tag}.json")).read().strip().splitlines()[-1])
        L += [f"bench.py's CUDA-event time for the same step, not under a profiler, is "
              f"{b['ms_per_step']} ms (profiles/bench_{tag}.json)."]
    except (OSError, ValueError, KeyError):
        pass
    open(os.path.join(PROF, f"launches_{tag}_summary.md"), "w").write("\n".join(L) + "\n")
    shutil.copyfile(path, os.path.join(PROF, f"launches_{tag}.csv"))     # the raw list as well


def raw_page(stem, tag):
    """The raw metric page of a capture: the CSV the GPU box exported, or the report."""
    path = os.path.join(OUT, f"{stem}_{tag}.raw.csv")
    if os.path.exists(path) and os.path.getsize(path):
        return open(path).read()
    rep = os.path.join(OUT, f"{stem}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        print("missing:", rep)
        return None
    return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                          text=True).stdout


def full(tag):
    raw = raw_page("decode_p1_4096", tag)
    if raw is None:
        return
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, d = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    name = d[idx["Kernel Name"]]
    L = ["# ncu --set full --clock-control none --import-source on -k regex:xa_decode -c 1 "
         "python bench.py --steps 1 --warmup 3 --no-extras",
         f"# kernel: {name}, BASELINE.json configs[1] (4096 mono 8-bit streams x 60 s), profile mix "
         f"P1, one B200; round {tag}", "metric,unit,value"]
    for m in METRICS:
        if m in idx:
            L.append(f"{m},{units[idx[m]]},{d[idx[m]]}")
    stalls = []
    for h in hdr:
        if "average_warp_latency_issue_stalled" in h or ("warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")):
            try:
                stalls.append((fnum(d[idx[h]]), h, units[idx[h]]))
            except ValueError:
                pass
    for v, h, u in sorted(stalls, reverse=True)[:8]:
        L.append(f"{h},{u},{v}")
    open(os.path.join(PROF, f"decode_p1_4096_{tag}_ncu_summary.csv"), "w").write("\n".join(L) + "\n")

    def gb(m):
        v, u = fnum(d[idx[m]]), units[idx[m]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    algo = 4096 * (2728704 + 5292000)
    json.dump({"kernel": name, "workload": "4096 mono 8-bit streams x 60 s, mix P1",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "algorithmic_bytes_per_launch": algo,
               "traffic_over_algorithmic": round((rd + wr) / algo, 4),
               "source": f"profiles/decode_p1_4096_{tag}_ncu_summary.csv (ncu --set full, one launch)"},
              open(os.path.join(PROF, "decode_traffic.json"), "w"), indent=1)


def secondary(tag, stem, cmd, what):
    """ncu --set full of one launch of a secondary kernel -> <stem>_<tag>_ncu_summary.csv"""
    raw = raw_page(stem, tag)
    if raw is None:
        return
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("empty:", stem)
        return
    hdr, units, d = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    L = [f"# ncu --set full --clock-control none --import-source on -c 1 python {cmd}",
         f"# kernel: {d[idx['Kernel Name']]}; {what}; one B200; round {tag}", "metric,unit,value"]
    for m in METRICS:
        if m in idx:
            L.append(f"{m},{units[idx[m]]},{d[idx[m]]}")
    stalls = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((fnum(d[idx[h]]), h, units[idx[h]]))
            except ValueError:
                pass
    for v, h, u in sorted(stalls, reverse=True)[:6]:
        L.append(f"{h},{u},{v}")
    open(os.path.join(PROF, f"{stem}_{tag}_ncu_summary.csv"), "w").write("\n".join(L) + "\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    secondary(tag, "decode_stereo8_p1",
              "tools/prof_decode.py --mix P1 --streams 2048 --seconds 30 --bits 8 --ch 2 --steps 1 --warmup 0",
              "stereo decode, direct form, 2048 stereo 8-bit streams x 30 s, mix P1")
    secondary(tag, "encode_stereo4",
              "tools/prof_decode.py --mix P0 --streams 2048 --seconds 30 --bits 4 --ch 2 --steps 1 --warmup 0 --encode",
              "reference-exact encode, 2048 stereo streams x 30 s -> 4-bit XA")
    secondary(tag, "search_stereo4",
              "tools/prof_decode.py --mix P0 --streams 1024 --seconds 4 --bits 4 --ch 2 --steps 1 --warmup 0 --search",
              "searching encoder (extension), 1024 stereo streams x 4 s -> 4-bit XA, 65 candidates per block")
    secondary(tag, "seg_mono8_p2",
              "tools/prof_decode.py --mix P2 --streams 2048 --seconds 30 --bits 8 --ch 1 --steps 1 --warmup 1",
              "segment form (xa_seg_kernel): every lane a fixed segment of one stream, all blocks through "
              "the chain step; 2048 mono 8-bit streams x 30 s, mix P2 (80 % chain blocks)")
    secondary(tag, "seg_stereo4_p2",
              "tools/prof_decode.py --mix P2 --streams 2048 --seconds 30 --bits 4 --ch 2 --steps 1 --warmup 1",
              "segment form, stereo: both channels side by side in every lane; 2048 stereo 4-bit streams x 30 s, mix P2")
    secondary(tag, "chain_mono8_p3",
              "tools/prof_decode.py --mix P3 --streams 4096 --seconds 4 --bits 8 --ch 1 --steps 1 --warmup 1",
              "chain form (xa_chain_kernel): a CTA per 32 streams, loader / stepper / storer warps; "
              "4096 mono 8-bit streams x 4 s without any cut block (latency-bound by design: one chain per stream)")
    secondary(tag, "chain_stereo8_p3",
              "tools/prof_decode.py --mix P3 --streams 4096 --seconds 4 --bits 8 --ch 2 --steps 1 --warmup 1",
              "chain form, stereo: a stepper warp per channel, the storer interleaves; "
              "4096 stereo 8-bit streams x 4 s without any cut block")
    secondary(tag, "relay_pass1_mono8_c20",
              "tools/prof_decode.py --mix C20 --streams 2048 --seconds 30 --bits 8 --ch 1 --steps 1 --warmup 1",
              "relay form, first pass: the direct form whose walker warps hand their stragglers on; "
              "2048 mono 8-bit streams x 30 s, 20 % chain blocks")
    secondary(tag, "relay_pass2_mono8_c20",
              "tools/prof_decode.py --mix C20 --streams 2048 --seconds 30 --bits 8 --ch 1 --steps 1 --warmup 1",
              "relay form, second pass: the dense walkers finishing the handed-on chains; same launch")
    for n in (f"bench_{tag}.json", f"bench_ref_{tag}.json", f"extras_{tag}.json", f"pcie_{tag}.json",
              f"latency_{tag}.json", f"prof_relay_{tag}.json", f"prof_seg_mono8_{tag}.json",
              f"prof_seg_stereo4_{tag}.json", f"prof_chain_mono8_{tag}.json", f"prof_chain_stereo8_{tag}.json", f"auto_sweep_{tag}.log"):
        copy(n)
    launches(tag)
    full(tag)


if __name__ == "__main__":
    main()
