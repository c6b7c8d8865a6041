#!/usr/bin/env python3
"""Host <-> device copy bandwidth of this box with pinned and pageable buffers (context for bench.py's e2e number)."""
import json
import sys

import torch

n = 300 * 2 ** 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for kind in ("pinned", "pageable"):
    host = torch.empty(n, dtype=torch.uint8)
    if kind == "pinned":
        host = host.pin_memory()
    for name, fn in (("h2d", lambda: dev.copy_(host, non_blocking=True)), ("d2h", lambda: host.copy_(dev, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out["%s_%s_GBs" % (kind, name)] = round(5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(out))
