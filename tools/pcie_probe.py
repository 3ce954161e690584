#!/usr/bin/env python
"""Host<->device copy rates of this box with pinned buffers: H2D alone, D2H alone
and both at once -- the ceiling the e2e leg of bench.py (host buffers in, host
buffers out) can reach.  Prints one JSON line."""
import json

import torch


def rate(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return nbytes * reps / (a.elapsed_time(b) * 1e6)


def main():
    n_in, n_out = 1397096448, 2709504000          # bytes per e2e step of bench.py
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    out = {"h2d_GBps": round(rate(h2d, n_in), 2), "d2h_GBps": round(rate(d2h, n_out), 2)}
    t = rate(both, n_in + n_out)
    out["duplex_GBps_total"] = round(t, 2)
    out["duplex_ms_per_e2e_step"] = round((n_in + n_out) / t / 1e6, 2)
    # the e2e metric if copies were the only cost: 512 streams x 2 646 000 samples a step
    out["e2e_ceiling_Msamples_s"] = round(512 * 2646000 / out["duplex_ms_per_e2e_step"] / 1e3, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
