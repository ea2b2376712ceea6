"""Print the metrics we track from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv
import re
import subprocess
import sys

PAT = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|"
    r"l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|"
    r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|"
    r"l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|st)\.sum|l1tex__t_requests_pipe_lsu_mem_global_op_(ld|st)\.sum|"
    r"lts__t_sectors_op_(read|write)\.sum|lts__t_sectors_srcunit_tex_op_(read|write)\.sum|"
    r"l1tex__throughput\.avg\.pct_of_peak_sustained_active|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_pipe_lsu_wavefronts\.sum|l1tex__lsu_writeback_active\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__thread_inst_executed_per_inst_executed\.ratio|smsp__inst_executed\.sum|sm__cycles_elapsed\.max|"
    r"smsp__cycles_active\.avg|local_load|local_store|smsp__inst_executed_op_local_(ld|st)\.sum)$")


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data:
        print("==", r[ki][:100])
        for i, h in enumerate(hdr):
            if PAT.search(h):
                v = r[i]
                if h.startswith("smsp__average_warps_issue_stalled") and float(v.replace(",", "") or 0) < 0.3:
                    continue
                print(f"  {h:86s} {v:>18s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
