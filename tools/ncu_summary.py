"""Per-launch summary of an ncu report (run here, no GPU needed):

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv ; python tools/ncu_summary.py /tmp/x.csv > profiles/x.json
"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_to_sm_read_sectors",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__inst_executed.sum": "warp_insts",
}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        rec = {"kernel": re.sub(r"\(.*", "", d["Kernel Name"])[:100]}
        for i, n in enumerate(names):
            if n in WANT:
                try:
                    rec[WANT[n] + ("_" + units[i].replace("/", "_per_") if units[i] else "")] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        out.append(rec)
    print(json.dumps({"source": sys.argv[1], "launches": out}, indent=1))


if __name__ == "__main__":
    main()
