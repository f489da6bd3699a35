"""Host-side packer throughput on this box: db200_hostpack over T Python threads (ctypes releases the GIL), page-locked and
pageable sources; plus a plain memcpy-like read (np.sum) for the memory-bandwidth context."""
import os, sys, time, json, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from dashing_b200 import capi

N = 1 << 30
rng = np.random.default_rng(0)
src_pageable = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=N)]
out = {"isa": capi.lib.db200_hostpack_isa().decode(), "cpus": len(os.sched_getaffinity(0)), "cpu_max": open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else None}
srcs = {"pageable": src_pageable}
try:
    pin = capi.pinned_empty(N); pin[:] = src_pageable; srcs["page_locked"] = pin
    dst_pin = capi.pinned_empty(N // 16 * 6)
except Exception as e:
    out["pin_error"] = repr(e); dst_pin = None
for name, src in srcs.items():
    for T in (1, 2, 4, 8, 12, 16, 24, 32):
        codes = np.zeros(N // 16, np.uint32) if dst_pin is None else dst_pin[: N // 16 * 4].view(np.uint32)
        valid = np.zeros(N // 16, np.uint16) if dst_pin is None else dst_pin[N // 16 * 4:].view(np.uint16)
        part = N // T // 4096 * 4096
        def work(i):
            b0 = i * part; b1 = N if i == T - 1 else b0 + part
            capi.lib.db200_hostpack(C.c_void_p(src.ctypes.data + b0), b1 - b0, C.cast(codes.ctypes.data + b0 // 16 * 4, C.POINTER(C.c_uint32)), C.cast(valid.ctypes.data + b0 // 16 * 2, C.POINTER(C.c_uint16)))
        best = 0
        for rep in range(3):
            th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
            t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
            best = max(best, N / dt / 1e9)
        out[f"{name}_T{T}_GBps"] = round(best, 1)
print(json.dumps(out))
