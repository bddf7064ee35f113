"""Static round-robin super-tile schedule: work imbalance over 148 CTAs (units = popcount of tile masks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from lidog_b200 import me as ME
from lidog_b200.lidog import synth
dev = torch.device("cuda", 0)
scans = synth.make_batch(8, 1234, "kitti", 7)
q = ME.utils.sparse_quantize_batch([torch.from_numpy(p).to(dev) for p, _ in scans], [torch.from_numpy(l).to(dev) for _, l in scans], 0.05, -1)
cm = ME.CoordinateManager.from_quantized(q)
for ts in (1, 2, 4, 8, 16):
    layer = ME.MinkowskiConvolution(32, 32, kernel_size=3, dimension=3)
    _, (pf, _, _, _) = layer._plans(cm, ts)
    m = pf.tile_mask.view(torch.int32).cpu().numpy().astype(np.uint32).reshape(-1)
    units = np.array([bin(int(v)).count("1") for v in m])
    for T in (1, 2, 4, 5, 8):
        ns = (len(units) + T - 1) // T
        su = np.add.reduceat(units, np.arange(0, len(units), T))
        per = np.zeros(148)
        for i, u in enumerate(su):
            per[i % 148] += u
        # dynamic (greedy in order)
        dyn = np.zeros(148)
        for u in su:
            dyn[np.argmin(dyn)] += u
        print(f"ts{ts} tiles={len(units)} T={T} super={ns} mean={per.mean():.1f} static max/mean={per.max()/per.mean():.3f} dynamic max/mean={dyn.max()/dyn.mean():.3f}")
