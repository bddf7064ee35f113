"""Host-side (Python) profile of the training step: where the CPU time of issuing one step goes.
    python tools/host_profile.py [--steps 3]
The step is GPU-bound only while the host can issue it faster than the GPU runs it (bench.py reports both:
ms_per_step and host_issue_ms_per_step)."""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from lidog_b200 import cabi
from lidog_b200.lidog import synth, model as M, step, bev as lbev
dev = torch.device("cuda", 0)
cabi.lib()
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234)
net = M.MinkUNet34BEV(1, 7, mapping_bound_2d=50.0).to(dev)
if lbev.CONFIG["channels_last"]:
    net.encoders2d.to(memory_format=torch.channels_last)
tr = step.LidogTrainer(net, num_classes=7, shape="kitti")
scans = synth.make_batch(8, 1234, "kitti", 7)
pts = [torch.from_numpy(p).to(dev) for p, _ in scans]
lab = [torch.from_numpy(l).to(dev) for _, l in scans]
for _ in range(3):
    tr.training_step(pts, lab)
torch.cuda.synchronize()
n = 3
t0 = time.perf_counter()
for _ in range(n):
    tr.training_step(pts, lab)
t_issue = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
t_total = (time.perf_counter() - t0) / n
print(f"host issue {1e3*t_issue:.1f} ms/step, wall {1e3*t_total:.1f} ms/step (no profiler)")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    tr.training_step(pts, lab)
pr.disable()
torch.cuda.synchronize()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[:48]))
