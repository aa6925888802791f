"""Condenses gpurun_out/prof_{va,hd}.ncu-rep (ncu --set full captures) into the tracked summaries under profiles/.

    python tools/profile_summary.py r01        # -> profiles/r01_va_ncu.txt, r01_hd_ncu.txt, traffic.json, ...
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "sm__cycles_elapsed.max.per_second"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def sass_mix(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    mix = {}
    for r in rows[2:]:
        if len(r) <= max(ie, isamp) or r[0].startswith("Kernel Name") or r[0] == "Address":
            if len(r) > 1 and r[0].startswith("Kernel Name"):
                break
            continue
        op = r[1].strip().split()
        o = (op[1] if op[0].startswith("@") and len(op) > 1 else op[0]).split(".")[0]
        m = mix.setdefault(o, [0, 0])
        m[0] += int(r[ie])
        m[1] += int(r[isamp])
    return mix


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    traffic = {}
    for name, stencil in (("va", "vert_adv"), ("hd", "hori_diff"), ("fused", None)):
        rep = "gpurun_out/prof_%s.ncu-rep" % name
        if not os.path.exists(rep):
            continue
        hdr, units, rows = raw(rep)
        lines = ["# ncu --set full --clock-control none, kernel regex %s_ (tools/gpu_round.sh ncu); one column per captured launch" % name]
        ik = hdr.index("Kernel Name")
        lines.append("kernel: " + rows[0][ik][:400])
        vals = {}
        for i, h in enumerate(hdr):
            if h in KEYS:
                vals[h] = [r[i] for r in rows]
                lines.append("%-84s %-10s %s" % (h, units[i], "  ".join(vals[h])))
        rd = [float(x) for x in vals["dram__bytes_read.sum"]]
        wr = [float(x) for x in vals["dram__bytes_write.sum"]]
        scale = 1e6 if units[hdr.index("dram__bytes_read.sum")] == "Mbyte" else 1.0
        per_launch = int((sum(rd) / len(rd) + sum(wr) / len(wr)) * scale)
        if stencil is not None:
            traffic[stencil] = per_launch
        lines.append("dram traffic per launch (read + write, mean of the captured launches): %d bytes" % per_launch)
        mix = sass_mix(rep)
        tot = sum(v[0] for v in mix.values()) or 1
        lines.append("")
        lines.append("# SASS opcode mix of the first captured launch (warp instructions executed, share, stall samples)")
        for o, (e, s) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:24]:
            lines.append("%-12s %10d %5.1f%% %6d" % (o, e, 100.0 * e / tot, s))
        open("profiles/%s_%s_ncu.txt" % (tag, name), "w").write("\n".join(lines) + "\n")
    if traffic:
        json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
    for src, dst in (("launches.csv", "%s_launches_va.csv" % tag), ("launches_hd.csv", "%s_launches_hd.csv" % tag),
                     ("tune.txt", "%s_variant_sweep.txt" % tag), ("bench.json", "%s_bench_va.json" % tag),
                     ("bench_hd.json", "%s_bench_hd.json" % tag), ("bench_ref.json", "%s_bench_reference.json" % tag),
                     ("gpu.txt", "%s_gpu.txt" % tag), ("pytest_gpu.log", "%s_pytest_gpu.log" % tag)):
        if os.path.exists("gpurun_out/" + src):
            shutil.copy("gpurun_out/" + src, "profiles/" + dst)


if __name__ == "__main__":
    main()
