"""What does a plain device copy reach when it moves as few bytes as the grid2d 1000^2 product (80 MB)?
MEASURED_PEAKS.json's HBM figure comes from large copies; a 12-20 us kernel never reaches that steady state.
Prints the time of torch's copy kernel (read + write = the bytes given) for a few sizes, warm (same buffers)
and cold (rotating through 8 buffer pairs, > L2)."""
import json
import torch

st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for mb in (40, 80, 160, 640):
        n = mb * 1024 * 1024 // 2 // 8          # half read, half written
        pairs = [(torch.rand(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")) for _ in range(8)]
        for mode in ("warm", "cold"):
            for i in range(10):
                a, b = pairs[i % 8 if mode == "cold" else 0]
                b.copy_(a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(st)
            for i in range(200):
                a, b = pairs[i % 8 if mode == "cold" else 0]
                b.copy_(a)
            e1.record(st)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 200 * 1e3
            print(json.dumps({"copy_MB_moved": mb, "cache": mode, "us": round(us, 2), "GBs": round(mb * 1.048576 / us * 1e3, 1)}), flush=True)
        del pairs
