#!/usr/bin/env python3
"""Extract the judged metrics from .ncu-rep files (ncu --set full) into one JSON:

    tools/ncu_summary.py out.json report1.ncu-rep [report2.ncu-rep ...]

Every kernel launch of every report becomes an entry keyed by the (shortened) kernel name."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"launch__(block_size|grid_size|registers_per_thread|occupancy_limit_registers|shared_mem_per_block_dynamic)|"
                  r"lts__t_sector_hit_rate\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__cycles_elapsed\.avg\.per_second|"
                  r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__sass_inst_executed_op_local_(ld|st)\.sum)$")


def main():
    out = {}
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        name_col = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[name_col]).replace("void ", "").replace("bemb::<unnamed>::", "").replace("<unnamed>::", "").strip()
            key, k = name, 1
            while key in out:
                k += 1
                key = f"{name} #{k}"
            ent = {"report": rep.split("/")[-1]}
            for i, h in enumerate(hdr):
                if KEEP.match(h):
                    ent[h] = {"value": r[i], "unit": units[i]}
            out[key] = ent
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    print(f"{len(out)} launches -> {sys.argv[1]}")


if __name__ == "__main__":
    main()
