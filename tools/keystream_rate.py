#!/usr/bin/env python
"""Throughput of the batched key-stream generator (keystream.generate): streams x steps per second."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gym_lorenz_b200 import keystream
for kind in ("lorenz4_pair", "memristive4_pair"):
    for n, num, burn in ((1, 6500, 1500), (4096, 6500, 1500), (65536, 3500, 1500), (1048576, 1564, 1500)):
        keystream.generate(burn + 64, n_streams=n, kind=kind, burn_in=burn)        # warm-up (allocator, module load)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        x = keystream.generate(num, n_streams=n, kind=kind, burn_in=burn)
        q = keystream.quantize(x)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"kind": kind, "streams": n, "steps": num, "burn_in": burn, "seconds": round(dt, 4),
                          "stream_steps_per_s": n * num / dt, "key_symbols": int(q.numel())}), flush=True)
        del x, q
