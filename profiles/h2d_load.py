#!/usr/bin/env python3
"""Micro-benchmark of the map load path (ccp4.parse -> HBM): where do the ~65 ms per 226 MB map go, and what does a
persistent page-locked staging ring buy?   usage: python profiles/h2d_load.py [n]"""
import io
import sys
import time

import numpy as np
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
nbytes = 4 * n ** 3
payload = np.random.default_rng(0).standard_normal(n ** 3, dtype=np.float32).tobytes()
torch.cuda.init()
torch.zeros(1, device="cuda")
torch.cuda.synchronize()


def t(label, fn, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
        del r
    print("%-70s %s ms" % (label, " ".join("%7.1f" % x for x in out)))


def a_pin_then_to():
    st = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    io.BytesIO(payload).readinto(memoryview(st.numpy()).cast("B"))
    return st.view(torch.float32).to("cuda", non_blocking=True), st


def b_empty_pinned():
    st = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    io.BytesIO(payload).readinto(memoryview(st.numpy()).cast("B"))
    return st.view(torch.float32).to("cuda", non_blocking=True), st


RING = None


def c_ring(chunk=32 << 20, slots=2):
    global RING
    if RING is None or RING[0].numel() != chunk or len(RING) != slots:
        RING = [torch.empty(chunk, dtype=torch.uint8, pin_memory=True) for _ in range(slots)]
    views = [memoryview(r.numpy()).cast("B") for r in RING]
    events = [None] * slots
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h = io.BytesIO(payload)
    off = k = 0
    while off < nbytes:
        s = k % slots
        if events[s] is not None:
            events[s].synchronize()
        m = min(chunk, nbytes - off)
        got = h.readinto(views[s][:m])
        dev[off:off + got].copy_(RING[s][:got], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        events[s] = ev
        off += got
        k += 1
    return dev.view(torch.float32)


def d_pageable():
    arr = np.frombuffer(io.BytesIO(payload).read(), dtype=np.float32)
    return torch.from_numpy(arr.copy()).to("cuda")


def e_parts():
    t0 = time.perf_counter()
    st = torch.empty(nbytes, dtype=torch.uint8)
    t1 = time.perf_counter()
    st = st.pin_memory()
    t2 = time.perf_counter()
    io.BytesIO(payload).readinto(memoryview(st.numpy()).cast("B"))
    t3 = time.perf_counter()
    d = st.view(torch.float32).to("cuda", non_blocking=True)
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    print("    parts: empty %.1f  pin_memory %.1f  readinto %.1f  to() call %.1f  sync %.1f ms; is_pinned=%s"
          % tuple([(b - a) * 1e3 for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))] + [st.is_pinned()]))
    return d, st


print("map of %d^3 = %.0f MB" % (n, nbytes / 1e6))
t("E  parts of the current path", e_parts)
t("A  torch.empty().pin_memory(); readinto; .to(non_blocking)  [current]", a_pin_then_to)
t("B  torch.empty(pin_memory=True); readinto; .to(non_blocking)", b_empty_pinned)
t("D  read() -> pageable copy -> .to()", d_pageable)
for chunk in (8, 16, 32, 64):
    for slots in (2, 3):
        t("C  ring %d x %d MB: chunked readinto + async copies" % (slots, chunk), lambda: c_ring(chunk << 20, slots), reps=4)
t("A  again", a_pin_then_to)
