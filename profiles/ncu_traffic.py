"""Reads `ncu --set full` reports (imported here with `ncu -i ... --page raw --csv`) and writes the per-launch figures
bench.py and the summaries cite: duration, DRAM bytes read / written, occupancy, issue activity, stall breakdown.

    python profiles/ncu_traffic.py TAG=path.ncu-rep [TAG=path ...] [--json out.json --key J6M6_B65536 --name random_step]

Prints one block per report; with --json merges {key: {name: read+write bytes}} into the traffic file bench.py reads."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__maximum_warps_per_active_cycle_pct": "theoretical_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__issue_inst0.avg.pct_of_peak_sustained_active": "no_issue_pct",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "launch__occupancy_limit_registers": "occ_limit_reg_blocks",
    "launch__occupancy_limit_warps": "occ_limit_warp_blocks",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "smsp__average_warp_latency_per_inst_issued.ratio": "cycles_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio": "stall_selected",
    "smsp__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_fp64.sum": "fp64_instructions",
}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    res = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {"kernel": r[names.index("Kernel Name")]}
        for k, short in WANT.items():
            if k in names:
                i = names.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if short == "time":
                    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
                    short_k = "time_us"
                elif short.startswith("dram_") and short != "dram_pct":
                    v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                    short_k = short + "_bytes"
                else:
                    short_k = short
                d[short_k] = v
        res.append(d)
    return res


def main():
    args = sys.argv[1:]
    jpath = key = None
    names = {}
    pos = []
    i = 0
    while i < len(args):
        if args[i] == "--json":
            jpath = args[i + 1]; i += 2
        elif args[i] == "--key":
            key = args[i + 1]; i += 2
        else:
            pos.append(args[i]); i += 1
    merged = {}
    for item in pos:
        tag, path = item.split("=", 1)
        for d in read(path):
            print("== %s: %s" % (tag, d.pop("kernel")[:150]))
            for k, v in d.items():
                print("   %-28s %s" % (k, ("%.4g" % v)))
            if "dram_read_bytes" in d:
                merged[tag] = int(d["dram_read_bytes"] + d["dram_write_bytes"])
                print("   %-28s %d" % ("dram_total_bytes", merged[tag]))
    if jpath and key:
        try:
            cur = json.load(open(jpath))
        except Exception:
            cur = {}
        cur.setdefault(key, {}).update(merged)
        json.dump(cur, open(jpath, "w"), indent=1)


if __name__ == "__main__":
    main()
