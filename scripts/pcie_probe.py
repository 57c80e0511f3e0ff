#!/usr/bin/env python
"""Developer probe: host<->device copy bandwidth (simplex and duplex) from pinned memory,
host topology.  python scripts/pcie_probe.py"""
import os, subprocess, time
import torch

def sh(cmd):
    try:
        print(f"$ {cmd}\n" + subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout)
    except Exception as e:
        print(cmd, "failed", e)

sh("nvidia-smi topo -m")
sh("lscpu | head -40")
sh("numactl -H")
sh("cat /proc/meminfo | head -5")
sh("nproc")
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
print("H2D GB/s", n / t(h2d) / 1e9)
print("D2H GB/s", n / t(d2h) / 1e9)
print("duplex GB/s (sum)", 2 * n / t(both) / 1e9)
# chunked (16 MB) copies
ck = 16 << 20
def h2d_ck():
    with torch.cuda.stream(s1):
        for o in range(0, n, ck): d1[o:o+ck].copy_(h1[o:o+ck], non_blocking=True)
print("H2D chunked 16MB GB/s", n / t(h2d_ck) / 1e9)
# host memcpy bandwidth single thread
import numpy as np
a = np.frombuffer(h1.numpy(), dtype=np.uint8); b = np.frombuffer(h2.numpy(), dtype=np.uint8)
t0 = time.perf_counter(); b[:] = a; print("host memcpy 1 thread GB/s", n / (time.perf_counter() - t0) / 1e9)
