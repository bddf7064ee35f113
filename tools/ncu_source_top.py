"""Top instructions of an `ncu --page source --csv` export, per kernel: executed warp instructions and stall samples.
    python tools/ncu_source_top.py gpurun_out/x_source.csv [kernel substring] [N]"""
import csv
import sys

path, pat, top = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else ""), int(sys.argv[3]) if len(sys.argv) > 3 else 25
kernel, hdr, rows = None, None, []


def flush():
    if kernel is None or pat not in kernel or not rows:
        return
    ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    src = hdr.index("Source")
    tot_i = sum(int(r[ie] or 0) for r in rows)
    tot_s = sum(int(r[ss] or 0) for r in rows)
    print(f"== {kernel[:110]}\n   {len(rows)} SASS instructions, {tot_i} warp instructions executed, {tot_s} samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {hdr[i]: sum(int(r[i] or 0) for r in rows) for i in stall_cols}
    print("   stalls:", ", ".join(f"{k[6:]} {100 * v / max(tot_s, 1):.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
    print("   -- by samples")
    for idx, r in sorted(enumerate(rows), key=lambda x: -int(x[1][ss] or 0))[:top]:
        st = max(stall_cols, key=lambda i: int(r[i] or 0))
        print(f"   {idx:5d} smp {100 * int(r[ss] or 0) / max(tot_s, 1):5.1f}% exe {100 * int(r[ie] or 0) / max(tot_i, 1):5.1f}% {hdr[st][6:]:12s} {r[src][:90]}")


for r in csv.reader(open(path)):
    if r and r[0] == "Kernel Name":
        flush()
        kernel, hdr, rows = r[1], None, []
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) >= len(hdr) - 2:
        rows.append(r)
flush()
