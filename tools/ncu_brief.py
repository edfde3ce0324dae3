"""Print a compact table of the metrics that matter from an ncu report: python tools/ncu_brief.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue act %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
        ("smsp__inst_executed.sum", "warp instr"), ("sm__cycles_elapsed.avg", "cycles"), ("smsp__cycles_active.avg", "smsp active cyc"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__occupancy_limit_registers", "occ lim regs"),
        ("l1tex__t_sector_hit_rate.pct", "l1 hit %"), ("lts__t_sector_hit_rate.pct", "l2 hit %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts")]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "barrier", "not_selected", "lg_throttle", "mio_throttle",
          "dispatch_stall", "no_instruction", "membar", "sleeping", "tex_throttle", "drain", "imc_miss", "branch_resolving", "selected"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen and "--all" not in sys.argv:
            continue
        seen.add(name)
        print("=====", name)
        for key, label in WANT:
            if key in idx:
                print(f"  {label:18s} {r[idx[key]]} {units[idx[key]]}")
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in idx:
                st.append((float(r[idx[k]].replace(",", "")), s))
        print("  stalls/issue: " + ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True) if v >= 0.05))


if __name__ == "__main__":
    main()
