#!/usr/bin/env python
"""Turn an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the short summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_divq.ncu-rep profiles/r01_divq_128cubed.md "title"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__warps_active.avg.per_cycle_active", "warps active / SMSP"),
    ("smsp__warps_eligible.avg.per_cycle_active", "warps eligible / SMSP"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction (of 32)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests (warp-level)"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1/TEX sector hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_bytes.sum.per_second", "L2 bytes/s"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    extra = sys.argv[4:]  # free-form lines appended (e.g. executed steps)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none); one section per profiled launch\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name', '?')[:100]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k, label in KEYS:
                if k in d:
                    f.write(f"| {label} (`{k}`) | {d[k]} | {u[k]} |\n")
            try:
                req = float(d["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
                sec = float(d["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
                f.write(f"| sectors / load request (derived) | {sec / req:.2f} | of 32 max |\n")
            except Exception:
                pass
            f.write("\n")
        for e in extra:
            f.write(e + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
