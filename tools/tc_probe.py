"""Probe the tcgen05 convolution kernels against the fp32 SIMT kernels on the GPU.

Each case runs in its own subprocess (a trapped kernel kills only that case):
    python tools/tc_probe.py            # all cases
    python tools/tc_probe.py one <gather> <fmt> <kind> <ts> <ksize> <cin> <cout>
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(gather, fmt, kind, ts_in, ksize, cin, cout):
    import numpy as np
    import torch
    import MinkowskiEngine as ME
    from lidog_b200.me import conv as meconv
    from tests.helpers import random_voxels
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    coords = random_voxels(rng, 9000, span=30)
    base = ME.SparseTensor(coordinates=torch.from_numpy(coords).to(dev), features=torch.ones(len(coords), 1, device=dev))
    cm = base.coordinate_manager
    n_in = cm.level(ts_in).n
    torch.manual_seed(0)
    cls = ME.MinkowskiConvolutionTranspose if kind == "up" else ME.MinkowskiConvolution
    layer = cls(cin, cout, kernel_size=ksize, stride=2 if kind in ("down", "up") else 1, dimension=3).to(dev)
    x0 = torch.randn(n_in, cin, device=dev).relu_()
    res = {}
    outs = {}
    for mode in ("off", fmt):
        meconv.CONFIG["tc"], meconv.CONFIG["gather"] = mode, gather
        x = x0.clone().requires_grad_(True)
        layer.kernel.grad = None
        y = layer(ME.SparseTensor(x, tensor_stride=ts_in, coordinate_manager=cm)).F
        torch.manual_seed(1)
        gy = torch.randn_like(y) * 1e-4
        y.backward(gy)
        torch.cuda.synchronize()
        outs[mode] = (y.detach().double(), x.grad.double(), layer.kernel.grad.double().clone())
    for name, a, b in zip(("y", "dx", "dw"), outs[fmt], outs["off"]):
        res[name] = float((a - b).norm() / b.norm().clamp_min(1e-30))
    print("RESULT " + json.dumps(res))


CASES = [
    ("same", 1, 3, 32, 32), ("same", 1, 3, 64, 64), ("same", 1, 3, 96, 96), ("same", 1, 3, 128, 96),
    ("same", 2, 3, 384, 256), ("same", 2, 3, 192, 128), ("down", 1, 2, 64, 64), ("up", 2, 2, 256, 128),
    ("identity", 1, 1, 128, 96),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        a = sys.argv[2:]
        one(int(a[0]), a[1], a[2], int(a[3]), int(a[4]), int(a[5]), int(a[6]))
        sys.exit(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    for gather in (0, 1):
        for fmt in ("fp16", "bf16"):
            for case in (CASES[:3] if quick else CASES):
                cmd = [sys.executable, __file__, "one", str(gather), fmt] + [str(c) for c in case]
                try:
                    r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
                    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
                    msg = line[0] if line else "FAIL rc=%d %s" % (r.returncode, (r.stderr.strip().splitlines() or ["?"])[-1][:300])
                except subprocess.TimeoutExpired:
                    msg = "TIMEOUT"
                print(f"gather={gather} fmt={fmt} case={case}: {msg}", flush=True)
