"""Host-side probe of the packed H2D path (runs on the GPU box; python scripts/pack_probe.py):
  A  pack rate into a large pinned buffer                 (what gnx_infer_host calibrates)
  B  pack rate into a small ring of pinned slots          (output stays in cache)
  C  A with a concurrent stream of raw pinned H2D copies  (DRAM / mesh contention)
  D  B with every slot copied to the device as soon as it is packed, raw copies on a second stream
Rates are GB/s of int8 input consumed."""
import os, sys, time, threading, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnomix_b200 import _lib
lib = _lib.lib()
Cn = 1_226_139
pitch = (Cn + 127) // 128 * 128
pw = pitch // 32
R = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
threads = lib.gnx_host_threads()
X = torch.empty((R, pitch), dtype=torch.int8).pin_memory()
X.random_(0, 2)
big = torch.empty((R, pw), dtype=torch.int64).pin_memory()
dev_pk = torch.empty((R, pw), dtype=torch.int64, device="cuda")
dev_x = torch.empty((512, pitch), dtype=torch.int8, device="cuda")
bad = C.c_int(0)
def pack(r0, n, out_ptr):
    lib.gnx_pack_rows_host(X.data_ptr() + r0 * pitch, n, pitch, Cn, out_ptr, pw, threads, C.byref(bad))
gb = lambda rows, t: rows * Cn / t * 1e-9
pack(0, 256, big.data_ptr())
# A
t0 = time.perf_counter(); pack(0, R, big.data_ptr()); tA = time.perf_counter() - t0
print("A  pack -> large buffer: %.1f GB/s (%d threads)" % (gb(R, tA), threads), flush=True)
# B
for blk in (16, 32, 64):
    ring = torch.empty((4, blk, pw), dtype=torch.int64).pin_memory()
    t0 = time.perf_counter()
    for i, r0 in enumerate(range(0, R - blk + 1, blk)):
        pack(r0, blk, ring[i % 4].data_ptr())
    tB = time.perf_counter() - t0
    print("B  pack -> ring of 4 x %d rows (%.1f MB per slot): %.1f GB/s" % (blk, blk * pw * 8e-6, gb(R // blk * blk, tB)), flush=True)
# raw H2D alone
s2 = torch.cuda.Stream()
def raw_copies(n_iter, stop):
    with torch.cuda.stream(s2):
        k = 0
        while not stop.is_set() and k < n_iter:
            dev_x.copy_(X[(k * 512) % (R - 512):(k * 512) % (R - 512) + 512], non_blocking=True)
            k += 1
            if k % 4 == 0:
                s2.synchronize()
    s2.synchronize()
    return k
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s2):
    e0.record(s2)
    for k in range(8):
        dev_x.copy_(X[k * 256:k * 256 + 512], non_blocking=True)
    e1.record(s2)
s2.synchronize()
print("   raw pinned H2D alone: %.1f GB/s" % (8 * 512 * pitch / (e0.elapsed_time(e1) * 1e-3) * 1e-9), flush=True)
# C
stop = threading.Event(); cnt = [0]
th = threading.Thread(target=lambda: cnt.__setitem__(0, raw_copies(10 ** 6, stop)))
th.start(); time.sleep(0.05)
t0 = time.perf_counter(); pack(0, R, big.data_ptr()); tC = time.perf_counter() - t0
stop.set(); th.join()
print("C  pack -> large buffer with concurrent raw H2D: pack %.1f GB/s, raw H2D %.1f GB/s (sum %.1f)" %
      (gb(R, tC), cnt[0] * 512 * pitch / tC * 1e-9, gb(R, tC) + cnt[0] * 512 * pitch / tC * 1e-9), flush=True)
# D
for blk in (16, 32, 64):
    ring = torch.empty((4, blk, pw), dtype=torch.int64).pin_memory()
    evs = [torch.cuda.Event() for _ in range(4)]
    s1 = torch.cuda.Stream()
    for with_raw in (False, True):
        stop = threading.Event(); cnt = [0]
        if with_raw:
            th = threading.Thread(target=lambda: cnt.__setitem__(0, raw_copies(10 ** 6, stop)))
            th.start(); time.sleep(0.05)
        used = [False] * 4
        t0 = time.perf_counter()
        for i, r0 in enumerate(range(0, R - blk + 1, blk)):
            sl = i % 4
            if used[sl]:
                evs[sl].synchronize()
            pack(r0, blk, ring[sl].data_ptr())
            with torch.cuda.stream(s1):
                dev_pk[r0:r0 + blk].copy_(ring[sl], non_blocking=True)
                evs[sl].record(s1)
            used[sl] = True
        s1.synchronize()
        tD = time.perf_counter() - t0
        if with_raw:
            stop.set(); th.join()
        rows = R // blk * blk
        raw = cnt[0] * 512 * pitch / tD * 1e-9
        print("D  ring of 4 x %d rows, slot copied when packed%s: pack %.1f GB/s%s" %
              (blk, " + concurrent raw H2D" if with_raw else "", gb(rows, tD), (", raw %.1f GB/s (sum %.1f)" % (raw, gb(rows, tD) + raw)) if with_raw else ""), flush=True)
