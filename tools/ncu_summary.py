#!/usr/bin/env python3
"""Summarise an ncu report (--set full) and a launch list (gpu__time_duration) into small text files for profiles/.
Usage: tools/ncu_summary.py <prof.ncu-rep> <launches.csv> <out prefix>"""
import csv, io, subprocess, sys, collections

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    rep, launches, prefix = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(prefix + "_full.txt", "w") as fh:
        fh.write(f"# from {rep} (ncu --set full --clock-control none), one launch per kernel\n")
        for r in rows[2:]:
            fh.write(f"\n## {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            for w in WANT:
                if w in hdr:
                    fh.write(f"{w:90s} {r[hdr.index(w)]} {units[hdr.index(w)]}\n")
    # DRAM traffic per launch of every captured kernel (bench.py copies the dominant kernel's figure into roofline.traffic)
    import json
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0]
        def val(metric):
            v = float(r[hdr.index(metric)].replace(",", ""))
            u = units[hdr.index(metric)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
        traffic[name.replace("_kernel", "")] = {"grid": r[hdr.index("Grid Size")], "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                                                 "duration_us_under_ncu": val("gpu__time_duration.sum")}
    with open(prefix + "_traffic.json", "w") as fh:
        json.dump({"source": rep, "how": "ncu --set full --clock-control none, one launch per kernel of the bench's all-CTUs-on pass (17 resident 4K pictures)", "kernels": traffic}, fh, indent=1)
    tot = collections.defaultdict(lambda: [0, 0.0])
    with open(launches) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].split("::")[-1]
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1000.0 if r["Metric Unit"] in ("nsecond", "ns") else v
        key = (name, r["Grid Size"])
        tot[key][0] += 1
        tot[key][1] += v
    with open(prefix + "_launches.txt", "w") as fh:
        fh.write(f"# from {launches} (ncu --metrics gpu__time_duration.sum --clock-control none); cold-cache serialised launches: compare SHARES\n")
        allus = sum(us for (_n, us) in tot.values())
        fh.write(f"{'kernel':28s} {'grid':18s} {'launches':>8s} {'avg us':>10s} {'share':>7s}\n")
        for (name, grid), (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{name:28s} {grid:18s} {n:8d} {us / n:10.1f} {us / allus:7.3f}\n")
    print(open(prefix + "_launches.txt").read())


if __name__ == "__main__":
    main()
