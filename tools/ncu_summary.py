"""Summarise an ncu report (one kernel launch per row) into a small JSON for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/eval_kernel_ncu_summary.json [kernel-regex]
"""
import csv
import json
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_busy_pct_when_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__cycles_active.avg": "sm_cycles_active_avg",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed_max",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum": "fadd",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum": "fmul",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum": "ffma",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not pat.search(name):
            continue
        k = {"kernel": name}
        for m, key in WANT.items():
            if m in hdr:
                v = r[hdr.index(m)]
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    continue
                u = units[hdr.index(m)]
                if key == "duration_us" and u == "ns":
                    v /= 1e3
                if key.startswith("dram_bytes") and u.lower().startswith("k"):
                    v *= 1e3
                if key.startswith("dram_bytes") and u.lower().startswith("m"):
                    v *= 1e6
                k[key] = v
        if "fadd" in k:
            k["executed_fp32_flop"] = k.get("fadd", 0) + k.get("fmul", 0) + 2 * k.get("ffma", 0)
            k["executed_fp64_flop"] = k.get("dadd", 0) + k.get("dmul", 0) + 2 * k.get("dfma", 0)
        if "dram_bytes_read" in k:
            k["dram_traffic_bytes"] = k.get("dram_bytes_read", 0) + k.get("dram_bytes_write", 0)
        kernels.append(k)
    json.dump({"source": rep, "note": "ncu --set full, cold cache, serialised replay: compare shares, not absolutes",
               "kernels": kernels}, open(out, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
