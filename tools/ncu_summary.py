"""Summarise an .ncu-rep (ncu --set full) per launch: the metrics the roofline argument needs.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<kernel>.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum(\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"gpu__time_duration\.sum|launch__(grid_size|block_size|registers_per_thread|occupancy_limit_.*)|"
                  r"lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second)?|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__inst_executed\.sum|sm__cycles_elapsed\.max|smsp__cycles_active\.avg|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|smsp__warp_issue_stalled_.*_per_warp_active\.pct)$")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    for r in data:
        print(f"Kernel Name = {r[name_col]}")
        for h, u, v in zip(hdr, units, r):
            if KEEP.match(h):
                print(f"{h} = {v} {u}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
