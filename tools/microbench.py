"""BASELINE configs[3]: voxelisation + kernel-map + single MinkowskiConvolution micro-benchmark sweep.

    python tools/microbench.py [--points 10000,100000,1000000,2000000] [--channels 32,64,96,128,256] [--reps 10]

One random-plane cloud per point count (synth.make_plane_cloud: 5-12 occupied neighbours per voxel at any size).
Per point count it prints JSON lines for
  * voxelisation: lg_quantize_points + ONE lg_coords_pyramid level (hash build, unique / inverse maps, colabels) --
    algorithmic bytes 16 N (points + labels) read, 8 N (inverse map) + 24 U written -> GB/s of the
    measured copy bandwidth (MEASURED_PEAKS.json);
  * kernel maps: the 3x3x3 map (natural order and mask-sorted), the stride-2 map and the transposed map --
    16 U_out read + 4 K n_slots written -> GB/s, and probes/s;
  * convolution: k3 stride 1, k2 stride 2, k2 stride 2 transposed for Cin = Cout in --channels: forward, dgrad, wgrad
    through the C ABI -> algorithmic TFLOP/s (2 pairs Cin Cout) of the measured sustained bf16 peak.
Times are CUDA events on the launching stream, `reps` launches after a warm-up launch; every input is far larger than
what a previous launch leaves in L2 only at >= 1 M points -- the small sizes are launch-latency measurements and are
labelled as such (`launch_bound`)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="10000,100000,1000000,2000000")
    ap.add_argument("--channels", default="32,64,96,128,256")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    from lidog_b200 import cabi
    from lidog_b200 import me as ME
    from lidog_b200.me import conv as meconv
    from lidog_b200.me import coords as mecoords
    from lidog_b200.lidog import synth

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"hbm_gbs": 6551.7, "bf16_tflops_sustained": 1377.7}
    hbm, tf = peaks["hbm_gbs"], peaks["bf16_tflops_sustained"]
    dev = torch.device("cuda", 0)
    L = cabi.lib()
    fmt, dt16 = cabi.FMT_FP16, torch.float16

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    for n_pts in [int(v) for v in args.points.split(",")]:
        pts_np, lab_np = synth.make_plane_cloud(n_pts, 1234)
        pts, lab = torch.from_numpy(pts_np).to(dev), torch.from_numpy(lab_np).to(dev)
        n = pts.shape[0]
        b = torch.zeros(n, dtype=torch.int32, device=dev)

        # ---- voxelisation (quantise + one hash level)
        def voxelise():
            q4 = ME.utils.quantize_points(pts, 0.05, b)
            return mecoords.build_levels(q4, labels=lab, ignore_label=-1, strides=(1,))[0]
        lv = voxelise()
        U = lv["n"]
        ms = timed(voxelise, args.reps)
        by = 16 * n + 8 * n + 24 * U
        emit(points=n, op="voxelise (quantise + hash + unique/inverse + colabels, incl. the one host sync)", ms=round(ms, 4),
             voxels=U, alg_mbytes=round(by / 1e6, 2), gbs=round(by / ms / 1e6, 1), frac_hbm=round(by / ms / 1e6 / hbm, 3),
             launch_bound=n < 500_000)
        q = ME.utils.sparse_quantize_batch([pts], [lab], 0.05, -1)
        cm = ME.CoordinateManager.from_quantized(q)

        # ---- kernel maps
        for kind, ts_in, ts_out, ks in (("same", 1, 1, 3), ("same_sorted", 1, 1, 3), ("down_sorted", 1, 2, 2), ("up", 2, 1, 2)):
            def build():
                cm.plans.pop((kind, ts_in, ts_out, ks), None)
                return cm.plan(kind, ts_in, ts_out, ks)
            p = build()
            ms = timed(build, args.reps)
            n_out = p.n_out
            by = 16 * n_out + 4 * p.K * p.n_slots if kind != "up" else 16 * n_out + 8 * p.n_slots
            emit(points=n, op=f"kernel map {kind} k{ks} ts{ts_in}->{ts_out}", ms=round(ms, 4), rows=n_out,
                 pairs=p.count_pairs(), alg_mbytes=round(by / 1e6, 2), gbs=round(by / ms / 1e6, 1),
                 frac_hbm=round(by / ms / 1e6 / hbm, 3), probes_per_s=round(p.K * n_out / ms * 1e3) if kind != "up" else None,
                 launch_bound=n < 500_000)

        # ---- single convolutions
        for C in [int(v) for v in args.channels.split(",")]:
            for kind, ts, ks in (("same", 1, 3), ("down", 1, 2), ("up", 2, 2)):
                cls = ME.MinkowskiConvolutionTranspose if kind == "up" else ME.MinkowskiConvolution
                layer = cls(C, C, kernel_size=ks, stride=2 if kind in ("down", "up") else 1, dimension=3)
                _, (p_fwd, p_dgrad, p_wgrad, flip) = layer._plans(cm, ts)
                K = ks ** 3
                pairs = p_fwd.count_pairs()
                x16 = torch.randn(p_fwd.n_in, C, device=dev).relu_().to(dt16)
                dy16 = torch.randn(p_fwd.n_out, C, device=dev).to(dt16)
                w16 = torch.randn(K, C, C, device=dev).to(dt16)
                w16t = w16.transpose(1, 2).contiguous()
                y = torch.empty(p_fwd.n_out, C, device=dev)
                dx = torch.empty(p_fwd.n_in, C, device=dev)
                dw = torch.empty(K, C, C, device=dev)
                s = cabi.stream()
                ops = {
                    "fwd": lambda: cabi.check(L.lg_conv_layer_forward(p_fwd.cref, x16.data_ptr(), C, None, C, w16.data_ptr(),
                                                                      w16t.data_ptr(), 0, fmt, None, y.data_ptr(), None, s)),
                    "dgrad": lambda: cabi.check(L.lg_conv_layer_backward(p_dgrad.cref, p_wgrad.cref, flip, x16.data_ptr(), C,
                                                                        dy16.data_ptr(), C, w16.data_ptr(), fmt, None,
                                                                        dx.data_ptr(), None, s)),
                    "wgrad": lambda: cabi.check(L.lg_conv_layer_backward(p_dgrad.cref, p_wgrad.cref, flip, x16.data_ptr(), C,
                                                                        dy16.data_ptr(), C, w16.data_ptr(), fmt, None,
                                                                        None, dw.data_ptr(), s)),
                }
                for name, fn in ops.items():
                    ms = timed(fn, args.reps)
                    fl = 2.0 * pairs * C * C
                    by = (p_fwd.n_in + p_fwd.n_out) * C * 2 + p_fwd.n_out * C * 4 + 8 * pairs  # 16-bit operands, fp32 result
                    emit(points=n, op=f"conv {kind} k{ks} ts{ts} {C}->{C} {name}", ms=round(ms, 4), rows_out=p_fwd.n_out,
                         pairs=pairs, tflops=round(fl / ms / 1e9, 1), frac_tensor=round(fl / ms / 1e9 / tf, 3),
                         gbs=round(by / ms / 1e6, 1), frac_hbm=round(by / ms / 1e6 / hbm, 3), launch_bound=n < 500_000)


if __name__ == "__main__":
    main()
