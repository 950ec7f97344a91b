#!/usr/bin/env python3
"""Phase breakdown of the persistent engine from a FLASHV_TRACE_FILE dump (clock64 cycles)."""
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
steps, grid, nw, pts = np.frombuffer(raw[:16], np.int32)
t = np.frombuffer(raw[16:], np.int64).reshape(steps, grid, nw, pts).astype(np.float64)
mhz = float(sys.argv[2]) if len(sys.argv) > 2 else 1965.0
us = lambda c: c / mhz
names = ["poll+stage delta", "stream+compute", "scan+exact", "publish", "end barrier"]
sel = slice(4, steps)  # skip the first steps (cold ring)
for w, wn in enumerate(["warp 0", "warp 1"]):
    print(wn)
    if pts >= 7:  # point 6 sits between the tensor-memory part and the ring part of the stream phase
        a_ = t[sel, :, w, 6] - t[sel, :, w, 1]
        b_ = t[sel, :, w, 2] - t[sel, :, w, 6]
        ok = (t[sel, :, w, 6] > 0) & (t[sel, :, w, 2] > 0)
        if ok.any():
            # even warps take the tensor-memory chunks first, odd warps (the last one) the shared-memory ones
            first, second = ("tensor memory", "shared memory") if w == 0 else ("shared memory", "tensor memory")
            print(f"   {'  1st: ' + first:20s} mean {us(a_[ok].mean()):7.2f} us")
            print(f"   {'  2nd: ' + second:20s} mean {us(b_[ok].mean()):7.2f} us")
    for p in range(5):
        d = t[sel, :, w, p + 1] - t[sel, :, w, p]
        ok = (t[sel, :, w, p + 1] > 0) & (t[sel, :, w, p] > 0)
        d = d[ok]
        if d.size:
            print(f"   {names[p]:20s} mean {us(d.mean()):7.2f} us   p10 {us(np.percentile(d, 10)):7.2f}   p90 {us(np.percentile(d, 90)):7.2f}   max {us(d.max()):7.2f}")
    full = t[sel, :, w, 5][1:] - t[sel, :, w, 5][:-1]
    full = full[(t[sel, :, w, 5][1:] > 0) & (t[sel, :, w, 5][:-1] > 0)]
    print(f"   {'whole step':20s} mean {us(full.mean()):7.2f} us")
