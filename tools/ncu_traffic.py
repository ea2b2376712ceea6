"""Write profiles/traffic_<workload>.json from ncu --set full captures of the two step kernels.

  python tools/ncu_traffic.py cfg5w gpurun_out/prof_lb_<tag>.ncu-rep gpurun_out/prof_mp_<tag>.ncu-rep

The file records dram__bytes_read.sum + dram__bytes_write.sum per launch together with a hash of the kernel
sources it was measured on; bench.py reports `roofline.traffic` only while that hash matches.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def traffic(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                          "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[-1]
    g = lambda name: float(r[hdr.index(name)].replace(",", "")) * UNIT.get(units[hdr.index(name)], 1.0)  # noqa: E731
    return r[hdr.index("Kernel Name")], g("dram__bytes_read.sum") + g("dram__bytes_write.sum"), r[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")]


def main():
    wl, rep_lb, rep_mp = sys.argv[1:4]
    k_lb, b_lb, t_lb = traffic(rep_lb)
    k_mp, b_mp, t_mp = traffic(rep_mp)
    out = {"lb_step_kernel_bytes_per_launch": b_lb, "mp_step_kernel_bytes_per_launch": b_mp,
           "lb_kernel": k_lb, "mp_kernel": k_mp, "ncu_duration": {"lb": t_lb, "mp": t_mp},
           "kernel_source_hash": bench.kernel_source_hash(),
           "source": f"ncu --set full --clock-control none ({os.path.basename(rep_lb)}, {os.path.basename(rep_mp)}): "
                     f"dram__bytes_read.sum + dram__bytes_write.sum of one launch, {wl}, N=1"}
    p = os.path.join(ROOT, "profiles", f"traffic_{wl}.json")
    json.dump(out, open(p, "w"), indent=1)
    print(open(p).read())


if __name__ == "__main__":
    main()
