#!/usr/bin/env python3
"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the small JSON summaries committed under profiles/.

usage: ncu_summary.py RAW.csv OUT.json [--traffic-key KEY --traffic-file profiles/dram_traffic.json --source TEXT]
Keeps the metrics profiles/README.md discusses (time, DRAM bytes, instruction and stall counts, launch shape)
of the FIRST kernel row, and can stamp that launch's DRAM traffic into dram_traffic.json together with the
device_abi of the library that was profiled (bench.py refuses a stamp taken with other device code).
"""
import csv, json, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "no_instruction", "math_pipe_throttle",
          "branch_resolving", "barrier", "mio_throttle", "lg_throttle", "dispatch_stall", "drain", "imc_miss", "membar",
          "sleeping", "tex_throttle", "selected"]
KEEP += [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in STALLS]


def main():
    raw, out = sys.argv[1], sys.argv[2]
    opts = dict(zip(sys.argv[3::2], sys.argv[4::2]))
    rows = [r for r in csv.reader(open(raw)) if r and not r[0].startswith("==")]
    names, units, first = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    res = {"kernel": first[col["Kernel Name"]]} if "Kernel Name" in col else {}
    for m in KEEP:
        if m in col:
            res[m] = {"value": first[col[m]], "unit": units[col[m]]}
    json.dump(res, open(out, "w"), indent=1)
    print(out, res.get("kernel"), res.get("gpu__time_duration.sum"))
    if "--traffic-key" in opts:
        def to_bytes(m):
            v, u = float(first[col[m]].replace(",", "")), units[col[m]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
        rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
        sys.path.insert(0, ".")
        from ad4cl_b200 import host
        path = opts.get("--traffic-file", "profiles/dram_traffic.json")
        try:
            tr = json.load(open(path))
        except Exception:
            tr = {}
        tr[opts["--traffic-key"]] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                     "device_abi": host.device_abi(),
                                     "source": opts.get("--source", out + " (ncu --set full, one launch)")}
        json.dump(tr, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
