"""Per-role wait-cycle breakdown of CTA 0 of k_gemm2 (LIDOG_DBG=8), one forward launch per case."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["LIDOG_DBG"] = str(int(os.environ.get("LIDOG_DBG", "0")) | 8)
import torch
from lidog_b200 import cabi
from lidog_b200 import me as ME
from lidog_b200.lidog import synth
dev = torch.device("cuda", 0)
scans = synth.make_batch(8, 1234, "kitti", 7)
q = ME.utils.sparse_quantize_batch([torch.from_numpy(p).to(dev) for p, _ in scans], [torch.from_numpy(l).to(dev) for _, l in scans], 0.05, -1)
cm = ME.CoordinateManager.from_quantized(q)
L = cabi.lib(); raw = C.CDLL(cabi.LIB_PATH)
names = ["prodA wait empty", "prodA issue", "prodB wait empty", "mma wait acc_empty", "mma wait fullB", "mma wait fullA", "mma issue+commit", "mma loop ovh", "epi wait acc_full", "epi store", "mma fences", "mma sched+prologue", "MMA ROLE LIFETIME", "units (CTA 0)", "super-tiles (CTA 0)", "prodA wait ids"]
for ts, cin, cout in ((1, 96, 96), (1, 32, 32), (8, 256, 256), (16, 256, 256)):
    layer = ME.MinkowskiConvolution(cin, cout, kernel_size=3, dimension=3)
    _, (pf, _, _, _) = layer._plans(cm, ts)
    x16 = torch.randn(pf.n_in, cin, device=dev).half(); w = torch.randn(27, cout, cin, device=dev).half(); y = torch.empty(pf.n_out, cout, device=dev)
    run = lambda: cabi.check(L.lg_conv_gemm_tc(pf.c, cabi.ptr(x16), cin, cabi.ptr(w), cout, 0, cabi.FMT_FP16, None, None, cabi.ptr(y), 2, cabi.stream()))
    run(); torch.cuda.synchronize()
    raw.lg_debug_profile(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    buf = (C.c_longlong * 16)(); raw.lg_debug_profile(buf, 1)
    tot = e0.elapsed_time(e1) * 1e-3 * 1.965e9
    print(f"ts{ts} {cin}->{cout}: {e0.elapsed_time(e1):.3f} ms = {tot:.0f} cycles")
    life = max(buf[12], 1)
    for n, v in zip(names, buf): print(f"   {n:20s} {v:10d} {100*v/tot:5.1f}% of event time, {100*v/life:5.1f}% of CTA-0 MMA-role lifetime")
