"""e2e (gnx_infer_host) throughput for pinned / pageable input and several packed fractions."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gnomix_b200 import synth, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
geom = synth.GEOMETRY["chr1"]
Cc, M, A, S, morgans = geom
W = Cc // M
base, smooth, (fx, fpop), _, _ = bench.build_models(geom)
ld = (Cc + 127) // 128 * 128
X = synth.admix_device(torch.from_numpy(fx).cuda(), n, morgans, seed=1, ld=ld)
Xpin = torch.empty((n, ld), dtype=torch.int8, pin_memory=True); Xpin.copy_(X)
Xpage = Xpin.numpy().copy()
Lh = torch.empty((n, W), dtype=torch.int32, pin_memory=True)
lib = _lib.lib(); hlr, hgbt = base.handle(), smooth.model.handle(S)
print("host threads", lib.gnx_host_threads(), flush=True)
def run(ptr, tag, steps=2, chunk=0):
    _lib.check(lib.gnx_infer_host(hlr, hgbt, ptr, n, ld, None, Lh.data_ptr(), chunk)); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(steps):
        _lib.check(lib.gnx_infer_host(hlr, hgbt, ptr, n, ld, None, Lh.data_ptr(), chunk))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / steps
    f, h, d = C.c_double(0), C.c_int64(0), C.c_int64(0)
    lib.gnx_infer_host_last_transfer(C.byref(f), C.byref(h), C.byref(d))
    print("%-28s %8.0f hap/s  %.3f s  frac=%.3f h2d=%.2f GB" % (tag, n / dt, dt, f.value, h.value / 1e9), flush=True)
    return Lh.clone()
os.environ["GNX_HOST_PACK"] = "0"
ref = run(Xpin.data_ptr(), "pinned raw")
run(Xpage.ctypes.data, "pageable raw", steps=1)
del os.environ["GNX_HOST_PACK"]
for th in ("", "8"):
    if th: os.environ["GNX_HOST_THREADS"] = th
    r = run(Xpin.data_ptr(), "pinned auto thr=%s" % (th or "all"))
    assert torch.equal(r, ref)
    pk, h2 = C.c_double(0), C.c_double(0); lib.gnx_infer_host_rates(C.byref(pk), C.byref(h2))
    print("   calibrated pack %.1f GB/s, h2d %.1f GB/s" % (pk.value, h2.value), flush=True)
os.environ.pop("GNX_HOST_THREADS", None)
for fr in ("0.3", "0.45", "0.6", "0.75", "0.9", "1"):
    os.environ["GNX_HOST_PACK_FRAC"] = fr
    assert torch.equal(run(Xpin.data_ptr(), "pinned frac=" + fr), ref)
os.environ["GNX_HOST_PACK_FRAC"] = "1"
assert torch.equal(run(Xpage.ctypes.data, "pageable frac=1"), ref)
del os.environ["GNX_HOST_PACK_FRAC"]
for ch in (256, 512, 768, 1024, 2048):
    assert torch.equal(run(Xpin.data_ptr(), "pinned auto chunk=%d" % ch, chunk=ch), ref)
    assert torch.equal(run(Xpage.ctypes.data, "pageable chunk=%d" % ch, chunk=ch), ref)
