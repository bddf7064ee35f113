"""Event trace of CTA 0 of k_gemm2 (LIDOG_DBG & 8): where a unit's time goes between the stage release, the
producer, and the MMA warp.  One forward launch of the block8 layer shape (ts 1, 96 -> 96).
    LIDOG_DBG=<extra bits> python tools/trace_units.py [ts cin cout]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["LIDOG_DBG"] = str(int(os.environ.get("LIDOG_DBG", "0")) | 8)
import numpy as np
import torch
from lidog_b200 import cabi
from lidog_b200 import me as ME
from lidog_b200.lidog import synth
ts, cin, cout = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1, 96, 96)
dev = torch.device("cuda", 0)
scans = synth.make_batch(8, 1234, "kitti", 7)
q = ME.utils.sparse_quantize_batch([torch.from_numpy(p).to(dev) for p, _ in scans], [torch.from_numpy(l).to(dev) for _, l in scans], 0.05, -1)
cm = ME.CoordinateManager.from_quantized(q)
L = cabi.lib(); raw = C.CDLL(cabi.LIB_PATH)
layer = ME.MinkowskiConvolution(cin, cout, kernel_size=3, dimension=3)
_, (pf, _, _, _) = layer._plans(cm, ts)
x16 = torch.randn(pf.n_in, cin, device=dev).half(); w = torch.randn(27, cout, cin, device=dev).half(); y = torch.empty(pf.n_out, cout, device=dev)
run = lambda: cabi.check(L.lg_conv_gemm_tc(pf.c, cabi.ptr(x16), cin, cabi.ptr(w), cout, 0, cabi.FMT_FP16, None, None, cabi.ptr(y), 2, cabi.stream()))
run(); torch.cuda.synchronize(); run(); torch.cuda.synchronize()
buf = (C.c_longlong * (5 * 512))(); raw.lg_debug_trace(buf)
t = np.array(buf[:], dtype=np.int64).reshape(5, 512)
free, arm, seen, commit, idc = t
n = int((commit > 0).sum())
print(f"ts{ts} {cin}->{cout}: {n} traced units of CTA 0")
t0 = idc[0]
lo, hi = (60, min(n, 120)) if n > 130 else (5, n)
print("unit   id_copy    p_free     p_arm    m_seen  m_commit | free-commit[u-sa..]  arm-free  seen-arm  commit-seen  period")
for u in range(lo, hi):
    rel = lambda a: int(a[u] - t0)
    best = min((int(free[u] - commit[v]) for v in range(max(0, u - 12), u) if free[u] >= commit[v]), default=-1)
    print(f"{u:4d} {rel(idc):9d} {rel(free):9d} {rel(arm):9d} {rel(seen):9d} {rel(commit):9d} | {best:9d} {int(arm[u]-free[u]):9d} {int(seen[u]-arm[u]):9d} {int(commit[u]-seen[u]):9d} {int(seen[u]-seen[u-1]):7d}")
d = lambda a, b: np.median((a - b)[lo:hi])
print("medians: arm-free", d(arm, free), " seen-arm", d(seen, arm), " commit-seen", d(commit, seen), " period", np.median(np.diff(seen[lo:hi])),
      " free-idcopy", d(free, idc))
