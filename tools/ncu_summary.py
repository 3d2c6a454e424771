"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers the
roofline discussion needs.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print(f"# {rep}: `ncu --set full --clock-control none --import-source on` (cold-cache, serialised replays)")
for r in rows[2:]:
    print(f"\nkernel: {r[hdr.index('Kernel Name')]}")
    print(f"grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
    for k in KEYS:
        if k in hdr:
            print(f"  {k:82s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    stalls = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr)
                     if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")), reverse=True)[:6]
    for v, h in stalls:
        print(f"  stall {h:76s} {v:18.2f} %")
