#!/usr/bin/env python
"""DRAM traffic per launch of the big kernels from an `ncu --set full` capture -> profiles/ncu_traffic_<workload>.json.

    python tools/ncu_traffic.py gpurun_out/r3_full.ncu-rep ml20m

bench.py puts the number of the dominant kernel into `roofline.traffic`, but only if the capture was taken from the kernel
sources as they are now: the file is stamped with a hash of rtrec_b200/csrc (bench.csrc_hash)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

GROUPS = {   # phase of bench.py -> kernels that make it up (first capture of each is used)
    "gram": ["gram_lower_kernel", "gram_head_tc_kernel", "gh_densify_kernel", "gram_mirror_kernel", "gram_unpermute_kernel",
             "gram_pull_cols_kernel", "row_sort_bitmap_kernel", "row_sort_small_kernel", "entry_pos_kernel"],
    "solve": ["slim_solve_warp_kernel", "slim_solve_kernel", "live_prefilter_kernel"],
    "recommend": ["recommend_tc_kernel", "recommend_tcfix_kernel", "recommend3_kernel", "recommend_sparse_kernel"],
}


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[0], rows[2:]
    k_i, r_i, w_i = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    t_i = hdr.index("gpu__time_duration.sum")
    units = rows[1]

    def to_bytes(v, unit):
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)

    first = {}
    for r in data:
        name = r[k_i]
        for key in sum(GROUPS.values(), []):
            if key in name and key not in first:
                first[key] = {"dram_bytes": to_bytes(r[r_i], units[r_i]) + to_bytes(r[w_i], units[w_i]), "duration": r[t_i] + " " + units[t_i]}
    out = {"csrc_sha16": bench.csrc_hash(), "source": os.path.basename(rep), "kernels": first}
    for g, ks in GROUPS.items():
        if any(k in first for k in ks):
            out[g] = int(sum(first[k]["dram_bytes"] for k in ks if k in first))
    path = os.path.join(ROOT, "profiles", f"ncu_traffic_{workload}.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
