"""Rebuilds profiles/traffic.json from `ncu --page raw --csv` exports of profiles/capture.sh (one launch of the dominant
kernel per workload) and copies the exports / launch lists to profiles/r2/.
usage: python profiles/update_traffic.py <tag prefix in gpurun_out, e.g. r2f>   (expects <prefix>_c2|c3|c4_raw.csv)"""
import csv, json, os, shutil, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

prefix = sys.argv[1] if len(sys.argv) > 1 else "r2f"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, ncu --set full --clock-control "
               "none (profiles/capture.sh); csrc_sha16 = bench.csrc_sha16() at capture time", "workloads": {}}
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for wl in ("C2", "C3", "C4"):
    src = os.path.join(root, "gpurun_out", f"{prefix}_{wl.lower()}_raw.csv")
    if not os.path.exists(src):
        continue
    rows = list(csv.reader(open(src)))
    hdr, units, first = rows[0], rows[1], rows[2]
    col = {k: i for i, k in enumerate(hdr)}
    rd = float(first[col["dram__bytes_read.sum"]]) * unit[units[col["dram__bytes_read.sum"]]]
    wr = float(first[col["dram__bytes_write.sum"]]) * unit[units[col["dram__bytes_write.sum"]]]
    dst = f"profiles/r2/r2_{wl.lower()}_ncu_raw.csv"
    shutil.copy(src, os.path.join(root, dst))
    ll = os.path.join(root, "gpurun_out", f"{prefix}_{wl.lower()}_launches.csv")
    if os.path.exists(ll):
        shutil.copy(ll, os.path.join(root, f"profiles/r2/r2_{wl.lower()}_launches.csv"))
    out["workloads"][wl] = {"kernel": first[col["Kernel Name"]], "grid": first[col["launch__grid_size"]],
                            "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                            "gpu_time_us": float(first[col["gpu__time_duration.sum"]]), "csrc_sha16": bench.csrc_sha16(),
                            "source": dst}
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out["workloads"], indent=1))
