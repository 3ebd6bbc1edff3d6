#!/usr/bin/env python
"""Raw PCIe rates of the box with pinned host memory (developer tool): H2D alone, D2H alone, both at once on two streams -
the ceiling of dgb_fv_step_host (bench.py `e2e`)."""
import json
import time

import torch


def main():
    n = 1 << 27                                  # 1 GiB of doubles
    h_in = torch.empty(n, dtype=torch.float64).pin_memory()
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    d_in = torch.empty(n, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return 8 * n * reps / (time.perf_counter() - t0) / 1e9
    for _ in range(2):
        run(True, True, 1)
    print(json.dumps({"h2d_alone_GBps": round(run(True, False), 1), "d2h_alone_GBps": round(run(False, True), 1),
                      "both_per_direction_GBps": round(run(True, True), 1)}))
    # chunked like the pipelined host path: 32 slabs per direction
    m = n // 32

    def chunked(reps=5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for k in range(32):
                with torch.cuda.stream(s1):
                    d_in[k * m:(k + 1) * m].copy_(h_in[k * m:(k + 1) * m], non_blocking=True)
                with torch.cuda.stream(s2):
                    h_out[k * m:(k + 1) * m].copy_(d_out[k * m:(k + 1) * m], non_blocking=True)
        torch.cuda.synchronize()
        return 8 * n * reps / (time.perf_counter() - t0) / 1e9
    print(json.dumps({"both_32_slabs_per_direction_GBps": round(chunked(), 1)}))


if __name__ == "__main__":
    main()
