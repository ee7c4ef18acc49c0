#!/usr/bin/env python3
"""Host-side probes for the numpy leg of the host pipeline: first-touch cost of fresh output arrays (page faults,
with and without transparent huge pages), float64 -> float32 conversion bandwidth by thread count.
python profiles/host_probe.py"""
import ctypes, mmap, os, sys, threading, time
import numpy as np
n = 172_800_000
def thp_info():
    for f in ("enabled", "defrag", "shmem_enabled"):
        try:
            print("thp", f, open(f"/sys/kernel/mm/transparent_hugepage/{f}").read().strip())
        except Exception as e:
            print("thp", f, e)
def anon_huge():
    for l in open("/proc/meminfo"):
        if l.startswith("AnonHugePages"):
            return l.strip()
thp_info()
def touch(arr, threads):
    flat = arr.reshape(-1).view(np.uint8)
    step = (flat.size + threads - 1) // threads
    def work(i):
        ctypes.memset(flat.ctypes.data + i * step, 1, min(step, flat.size - i * step))
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    return time.perf_counter() - t0
for threads in (1, 4, 8, 16):
    a = np.empty((3, n), dtype=np.float32)
    dt = touch(a, threads)
    print(f"np.empty first touch, {threads:2d} threads: {dt*1e3:7.1f} ms   {anon_huge()}")
    dt2 = touch(a, threads)
    print(f"          second touch, {threads:2d} threads: {dt2*1e3:7.1f} ms")
    del a
for threads in (1, 8, 16):
    m = mmap.mmap(-1, 12 * n + (2 << 20))
    m.madvise(mmap.MADV_HUGEPAGE)
    a = np.frombuffer(m, dtype=np.float32, count=3 * n)
    a = a.reshape(3, n)
    dt = touch(a, threads)
    print(f"mmap+MADV_HUGEPAGE first touch, {threads:2d} threads: {dt*1e3:7.1f} ms   {anon_huge()}")
    del a; m = None
for threads in (1, 8, 16):
    m = mmap.mmap(-1, 12 * n + (2 << 20))
    m.madvise(mmap.MADV_NOHUGEPAGE)
    a = np.frombuffer(m, dtype=np.float32, count=3 * n).reshape(3, n)
    dt = touch(a, threads)
    print(f"mmap+MADV_NOHUGEPAGE first touch, {threads:2d} threads: {dt*1e3:7.1f} ms   {anon_huge()}")
    del a; m = None
