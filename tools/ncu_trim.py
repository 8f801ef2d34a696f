"""Trim an `ncu -i X.ncu-rep --page raw --csv` dump to the metrics the roofline discussion uses, transposed
(one metric per line, one column per captured launch).  usage: python tools/ncu_trim.py raw.csv out.csv"""
import csv
import re
import sys

KEEP = re.compile(r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|"
                  r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(ld|st))?\.sum|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"launch__(registers_per_thread|shared_mem_per_block_dynamic|cluster_size|grid_size|block_size|waves_per_multiprocessor|"
                  r"occupancy_limit_\w+|cluster_max_active)|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__cycles_elapsed\.max|sm__cycles_active\.avg|smsp__inst_executed\.sum|sm__inst_executed_pipe_\w+\.sum|"
                  r"sm__pipe_tensor\w*cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"sm__inst_executed_pipe_tensor\w*\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps?_\w*issue_stalled_\w+_per_issue_active\.ratio|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__cycles_active\.avg|sm__sass_inst_executed_op_shared_(ld|st)\.sum)$")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = csv.writer(open(sys.argv[2], "w", newline=""))
    out.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
    for i, h in enumerate(hdr):
        if KEEP.match(h):
            out.writerow([h, units[i]] + [r[i] for r in data])


if __name__ == "__main__":
    main()
