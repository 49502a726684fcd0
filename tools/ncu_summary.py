"""Summarises one kernel of an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [row] [out.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "kernel": "Kernel Name", "duration_ms": "gpu__time_duration.sum", "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size", "block": "launch__block_size", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "fp64_cycles_active_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_active": "smsp__issue_active.avg.per_cycle_active", "inst_executed": "smsp__inst_executed.sum",
    "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dfma_thread_per_cycle_smsp": "smsp__sass_thread_inst_executed_op_dfma_pred_on.avg.per_cycle_elapsed",
    "dmul_thread_per_cycle_smsp": "smsp__sass_thread_inst_executed_op_dmul_pred_on.avg.per_cycle_elapsed",
    "dadd_thread_per_cycle_smsp": "smsp__sass_thread_inst_executed_op_dadd_pred_on.avg.per_cycle_elapsed",
    "local_loads": "smsp__sass_inst_executed_op_local_ld.sum", "local_stores": "smsp__sass_inst_executed_op_local_st.sum",
    "shared_loads": "sass__inst_executed_shared_loads", "shared_stores": "sass__inst_executed_shared_stores",
    "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
    "sm_cycles": "sm__cycles_elapsed.avg", "sm_ghz": "sm__cycles_elapsed.avg.per_second",
}


def main():
    rep = sys.argv[1]
    row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2 + row]))
    u = dict(zip(hdr, units))
    res = {}
    for k, name in KEYS.items():
        if name in d:
            v = d[name]
            try:
                v = float(v.replace(",", ""))
                if u.get(name) in ("us", "usecond") and k == "duration_ms":
                    v /= 1e3
                elif u.get(name) in ("ns", "nsecond") and k == "duration_ms":
                    v /= 1e6
                elif u.get(name) in ("s", "second") and k == "duration_ms":
                    v *= 1e3
            except ValueError:
                pass
            res[k] = v
    stalls = {}
    for name in hdr:
        if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
            try:
                stalls[name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(d[name])
            except ValueError:
                pass
    res["stall_cycles_per_issue"] = {k: v for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v >= 0.005}
    res["duration_unit_in_report"] = u.get("gpu__time_duration.sum")
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
