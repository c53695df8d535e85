#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters the
design is argued from: duration, DRAM bytes, issue-slot / pipe utilisation, instruction counts per
warp-bin-step, shared-memory bank conflicts and the warp-stall breakdown.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [--bins-per-launch N] > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_cbu.sum",
    "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    bins = None
    if "--bins-per-launch" in sys.argv:
        bins = float(sys.argv[sys.argv.index("--bins-per-launch") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}  (ncu --set full --clock-control none; per launch)")
    for r in rows[2:]:
        get = lambda k: r[hdr.index(k)] if k in hdr else None
        print("\n== " + get("Kernel Name")[:110])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:72s} {get(k):>18s} {units[hdr.index(k)]}")
        st = [(float(r[i]), hdr[i][len(STALL):].replace("_per_issue_active.ratio", "")) for i in range(len(hdr))
              if hdr[i].startswith(STALL) and hdr[i].endswith("_per_issue_active.ratio") and r[i]]
        st.sort(reverse=True)
        print("  warp stalls per issued instruction (top 8): " + ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
        if bins:
            inst = float(get("smsp__inst_executed.sum"))
            dur = float(get("gpu__time_duration.sum"))
            unit = units[hdr.index("gpu__time_duration.sum")]
            scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)
            print(f"  derived: {inst / (bins / 32):.1f} SASS warp-instructions per warp-bin-step; "
                  f"{bins / (dur * scale) / 1e9:.1f} Gbins/s under the profiler (not a bench value)")


if __name__ == "__main__":
    main()
