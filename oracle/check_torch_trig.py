#!/usr/bin/env python
"""Exhaustive pin of oracle/torch_trig.h against torch.sin / torch.cos (CPU fp32): EVERY float with |x| <= 10000,
both signs, zeros and denormals included.  Run in the build container (needs nothing but torch); takes minutes.
Writes oracle/check_torch_trig.log.  TEST INFRASTRUCTURE."""
import ctypes, os, sys, time
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import cpu_oracle as O  # noqa: E402

lib = O.lib()
top = int(np.float32(10000.0).view(np.uint32))          # bit pattern of 10000.0f
CH = 1 << 26
bad_s = bad_c = total = 0
t0 = time.time()
for sign in (0, 0x80000000):
    for lo in range(0, top + 1, CH):
        hi = min(lo + CH, top + 1)
        bits = (np.arange(lo, hi, dtype=np.uint32) | np.uint32(sign))
        x = bits.view(np.float32)
        s, c = np.empty_like(x), np.empty_like(x)
        # the oracle's scalar loop, split over threads by hand (ctypes releases the GIL)
        from concurrent.futures import ThreadPoolExecutor
        n = x.size
        parts = [(i * n // 8, (i + 1) * n // 8) for i in range(8)]
        def run(p):
            a, b = p
            lib.dvfo_torch_trig(x[a:b].ctypes.data_as(ctypes.c_void_p), ctypes.c_long(b - a),
                                s[a:b].ctypes.data_as(ctypes.c_void_p), c[a:b].ctypes.data_as(ctypes.c_void_p))
        with ThreadPoolExecutor(8) as ex:
            list(ex.map(run, parts))
        xt = torch.from_numpy(x)
        ts, tc = torch.sin(xt).numpy(), torch.cos(xt).numpy()
        bad_s += int((s.view(np.uint32) != ts.view(np.uint32)).sum())
        bad_c += int((c.view(np.uint32) != tc.view(np.uint32)).sum())
        total += n
msg = (f"torch {torch.__version__} CPU ({torch.backends.cpu.get_cpu_capability()}): {total} fp32 values with |x| <= 10000 "
       f"(every bit pattern, both signs): sin mismatches {bad_s}, cos mismatches {bad_c}; {time.time() - t0:.0f} s\n")
print(msg, end="")
open(os.path.join(HERE, "check_torch_trig.log"), "w").write(msg)
sys.exit(0 if bad_s == 0 and bad_c == 0 else 1)
