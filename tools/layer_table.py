"""Per-layer time table of the sparse-convolution launches of ONE training step (CUDA events per launch).

    python tools/layer_table.py [--batch 8] [--shape kitti] [--steps 3]

Groups the launches of `lidog_b200.me.conv` by (kernel, plan kind, tensor stride, Cin, Cout) and prints, per
group: launches per step, ms per step, algorithmic TFLOP/s (2 * pairs * Cin * Cout).  This is what decides
which layer shapes the tuning of csrc/conv_tc2.cu goes after; the totals match bench.py's `kernels` object.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--shape", default="kitti")
    ap.add_argument("--classes", type=int, default=7)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()

    from lidog_b200 import cabi
    from lidog_b200.lidog import synth, model as M, step, bev as lbev
    from lidog_b200.me import conv as meconv

    dev = torch.device("cuda", 0)
    cabi.lib()
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)
    cfg = synth.SHAPES[args.shape]
    net = M.MinkUNet34BEV(1, args.classes, mapping_bound_2d=cfg["bound"]).to(dev)
    if lbev.CONFIG["channels_last"]:
        net.encoders2d.to(memory_format=torch.channels_last)
    tr = step.LidogTrainer(net, num_classes=args.classes, shape=args.shape)
    scans = synth.make_batch(args.batch, 1234, args.shape, args.classes)
    pts = [torch.from_numpy(p).to(dev) for p, _ in scans]
    lab = [torch.from_numpy(l).to(dev) for _, l in scans]
    for _ in range(3):
        tr.training_step(pts, lab)
    meconv.PROFILE.update(enabled=True, events=False)
    meconv.PROFILE["records"].clear()
    tr.training_step(pts, lab)
    torch.cuda.synchronize()
    meconv.PROFILE.update(enabled=True, events=True)
    meconv.PROFILE["records"].clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.training_step(pts, lab)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / args.steps
    groups = {}
    for r in meconv.PROFILE["records"]:
        key = (r["kernel"], str(r["key"]), r["cin"], r["cout"])
        g = groups.setdefault(key, dict(ms=0.0, flops=0.0, n=0))
        g["ms"] += r["e0"].elapsed_time(r["e1"])
        g["flops"] += r["flops"]
        g["n"] += 1
    meconv.PROFILE.update(enabled=False, events=False)
    rows = []
    for (kern, key, cin, cout), g in groups.items():
        rows.append(dict(kernel=kern, plan=key, cin=cin, cout=cout, launches_per_step=g["n"] / args.steps,
                         ms_per_step=round(g["ms"] / args.steps, 4),
                         ms_per_launch=round(g["ms"] / g["n"], 4),
                         tflops=round(g["flops"] / (g["ms"] * 1e-3) / 1e12, 1) if g["ms"] else None))
    rows.sort(key=lambda r: -r["ms_per_step"])
    tot = {}
    for r in rows:
        tot[r["kernel"]] = tot.get(r["kernel"], 0.0) + r["ms_per_step"]
    print(json.dumps(dict(step_ms=round(step_ms, 3), totals_ms={k: round(v, 3) for k, v in tot.items()})))
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
