"""Summarise an `ncu --set full` report (run HERE, no GPU needed: `ncu -i <rep> --page raw --csv`) into a markdown table and a
small JSON file that bench.py reads for `roofline.traffic`.

    python profiles/ncu_summary.py gpurun_out/r02_gemm_syrk.ncu-rep profiles/r02_ncu_gemm_syrk.md [--json-kernel gemm_tn_tma_kernel profiles/r02_ncu_gemm_tn_tma.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe inst %"),
    ("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_active", "FP64 tensor path %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA (FP64 tensor) inst % of peak"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe cycles active %"),
    ("SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "DMMA sub-pipe cycles active (avg / SMSP)"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall: barrier %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    return header, units, data


def main():
    rep, md = sys.argv[1], sys.argv[2]
    header, units, data = load(rep)
    col = {h: i for i, h in enumerate(header)}
    name_col = col.get("Kernel Name", col.get("Function Name"))
    lines = ["# ncu --set full summary of `%s`" % rep.split("/")[-1], "",
             "Captured under gpurun on a B200 (`--clock-control none`); durations are serialised, cold-cache single launches: compare shares,",
             "not absolutes (profiles/README.md). One column per captured launch.", ""]
    names = [r[name_col].split("(")[0].split("::")[-1] for r in data]
    lines.append("| metric | unit | " + " | ".join("%s #%d" % (n[:28], i) for i, n in enumerate(names)) + " |")
    lines.append("|---|---|" + "---|" * len(names))
    summary = []
    for key, label in KEYS:
        if key not in col:
            continue
        lines.append("| %s (`%s`) | %s | " % (label, key, units[col[key]]) + " | ".join(r[col[key]] for r in data) + " |")
    open(md, "w").write("\n".join(lines) + "\n")
    if "--json-kernel" in sys.argv:
        k = sys.argv[sys.argv.index("--json-kernel") + 1]
        path = sys.argv[sys.argv.index("--json-kernel") + 2]
        for r in data:
            if k in r[name_col]:
                def num(key):
                    v = float(r[col[key]].replace(",", ""))
                    u = units[col[key]].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1,
                                "second": 1e9}.get(u, 1)
                d = {"kernel": r[name_col].split("(")[0], "report": rep.split("/")[-1], "grid": r[col["launch__grid_size"]],
                     "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                     "duration_ns": num("gpu__time_duration.sum")}
                for key in ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
                            "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
                    if key in col:
                        d[key] = float(r[col[key]].replace(",", ""))
                json.dump(d, open(path, "w"), indent=1)
                break
    print("wrote", md)


if __name__ == "__main__":
    main()
