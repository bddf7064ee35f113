import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import MinkowskiEngine as ME
from lidog_b200.me import norm
from tests.helpers import random_voxels
cuda = torch.device('cuda')
coords = random_voxels(np.random.default_rng(3), 5000, span=40)
base = ME.SparseTensor(coordinates=torch.from_numpy(coords).to(cuda), features=torch.ones(coords.shape[0], 1, device=cuda))
cm = base.coordinate_manager; n = coords.shape[0]
def rel(a, b):
    b = b.double(); return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
for C in (96, 128, 192, 256):
    torch.manual_seed(C)
    x = (torch.randn(n, C, device=cuda) * 2 + 0.5)
    gy = torch.randn(n, C, device=cuda) * 1e-3
    res = {}
    for fused in (1, 0):
        norm.CONFIG["fused"] = fused
        a = ME.MinkowskiBatchNorm(C).to(cuda)
        with torch.no_grad():
            g = torch.Generator(device="cpu").manual_seed(1)
            a.bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
            a.bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
        xs = x.clone().requires_grad_(True)
        y = ME.MinkowskiReLU()(a(ME.SparseTensor(xs, coordinate_manager=cm))).F
        y.backward(gy)
        res[fused] = (y.detach(), xs.grad, a.bn.weight.grad, a.bn.bias.grad)
    # float64 reference
    xd = x.double().requires_grad_(True)
    bn = torch.nn.BatchNorm1d(C).to(cuda).double()
    with torch.no_grad():
        bn.weight.copy_(a.bn.weight.double()); bn.bias.copy_(a.bn.bias.double())
    yd = torch.relu(bn(xd)); yd.backward(gy.double())
    ref = (yd.detach(), xd.grad, bn.weight.grad, bn.bias.grad)
    print(C, "fused-vs-torch", [f"{rel(p, q):.2e}" for p, q in zip(res[1], res[0])])
    print(C, "fused-vs-f64  ", [f"{rel(p, q):.2e}" for p, q in zip(res[1], ref)])
    print(C, "torch-vs-f64  ", [f"{rel(p, q):.2e}" for p, q in zip(res[0], ref)])
    d = (res[1][1].double() - ref[1]).abs()
    print("   worst channel of fused dx:", int(d.max(0).values.argmax()), float(d.max()), "gamma there", float(a.bn.weight[int(d.max(0).values.argmax())]))
