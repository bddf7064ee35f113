"""Per-kernel HBM summary of an `ncu --page raw` text export made by tools/gpu_run.sh (ncu_hbm): for every kernel
name, launches, total time, DRAM bytes moved and the achieved fraction of the measured copy bandwidth
(MEASURED_PEAKS.json).  Times under ncu are cold-cache and serialised: fractions are per-kernel statements, the
share of the step comes from the launch list."""
import collections
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main(path):
    peak = 6551.7
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        d = json.load(open(pk))
        for k, v in d.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                peak = float(v)
                break
    rows = collections.OrderedDict()
    cur = None
    for line in open(path):
        line = line.strip()
        if line.startswith("Kernel Name ="):
            name = line.split("=", 1)[1].strip()
            name = re.sub(r"\(.*$", "", name).replace("void ", "").strip()
            cur = rows.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "inst": 0.0})
            cur["n"] += 1
            continue
        m = re.match(r"(\S+) = ([0-9.eE+-]+) ?(\S*)", line)
        if not m or cur is None:
            continue
        key, val, unit = m.group(1), float(m.group(2)), m.group(3)
        if key == "dram__bytes_read.sum":
            cur["rd"] += val * UNIT.get(unit, 1.0)
        elif key == "dram__bytes_write.sum":
            cur["wr"] += val * UNIT.get(unit, 1.0)
        elif key == "gpu__time_duration.sum":
            cur["t"] += val * TIME.get(unit, 1e-9)
        elif key == "smsp__inst_executed.sum":
            cur["inst"] += val
    print(f"# {os.path.basename(path)}; peak = {peak} GB/s (measured copy bandwidth)")
    print(f"{'kernel':44s} {'n':>4s} {'ms':>8s} {'MB rd':>9s} {'MB wr':>9s} {'GB/s':>8s} {'frac':>6s} {'Minst':>8s}")
    for name, r in sorted(rows.items(), key=lambda kv: -kv[1]["t"]):
        if r["t"] <= 0:
            continue
        bw = (r["rd"] + r["wr"]) / r["t"] / 1e9
        print(f"{name[:44]:44s} {r['n']:4d} {r['t'] * 1e3:8.3f} {r['rd'] / 1e6:9.1f} {r['wr'] / 1e6:9.1f} {bw:8.0f} {bw / peak:6.2f} "
              f"{r['inst'] / 1e6:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
