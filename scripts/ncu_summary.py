"""Key metrics of every kernel in an .ncu-rep (`ncu --set full`), as text for profiles/: duration, DRAM bytes and GB/s, tensor / shared /
XU / issue pipe utilisation, the top warp stall reasons.  usage: ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
col = {k: i for i, k in enumerate(H)}


def get(r, k, default=float("nan")):
    try:
        return float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError):
        return default


UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("grb::", "")
    t = get(r, "gpu__time_duration.sum") * TIME.get(U[col["gpu__time_duration.sum"]], 1.0)
    rd = get(r, "dram__bytes_read.sum") * UNIT.get(U[col["dram__bytes_read.sum"]], 1)
    wr = get(r, "dram__bytes_write.sum") * UNIT.get(U[col["dram__bytes_write.sum"]], 1)
    print(f"== {name}   grid {r[col['Grid Size']]} block {r[col['Block Size']]}  regs {r[col['launch__registers_per_thread']]}")
    print(f"   duration {t:9.1f} us   DRAM read {rd / 1e6:9.1f} MB  write {wr / 1e6:9.1f} MB  -> {(rd + wr) / t / 1e3:7.1f} GB/s")
    for label, k in (("tensor pipe active (sm__pipe_tc_cycles_active)", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("tensor pipe active (sm__pipe_tensor_cycles_active)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("tensor instr (sm__inst_executed_pipe_tensor)", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
                     ("shared-memory pipe active", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("XU (MUFU) pipe", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                     ("FMA pipe", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                     ("ALU pipe", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     ("issue slots busy", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     ("warps active (occupancy)", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                     ("DRAM throughput", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     ("L2 throughput", "lts__throughput.avg.pct_of_peak_sustained_elapsed")):
        v = get(r, k)
        if v == v:
            print(f"   {label:52s} {v:6.1f} %")
    stalls = [(get(r, k), k.split("issue_stalled_")[1].split("_per_")[0]) for k in H
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    stalls = sorted([s for s in stalls if s[0] == s[0]], reverse=True)[:6]
    print("   warp stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls))
