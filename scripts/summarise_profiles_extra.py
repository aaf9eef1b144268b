"""Summaries of the late round-2 captures (gpurun_out/p3_*): the paired visit of the partitioned LU, the split residual of batched
small scenes, and the launch shares of a C3 sweep.  usage: python scripts/summarise_profiles_extra.py"""
import csv, io, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def summarise(rep, out, header):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units = rr[0], rr[1]
    with open(os.path.join(P, out), "w") as fh:
        fh.write(header + "\n\n")
        for r in rr[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            fh.write("kernel: %s  grid %s block %s\n" % (d["Kernel Name"][:100], d["Grid Size"], d["Block Size"]))
            for k in KEYS:
                if k in d:
                    fh.write("  %-75s %s %s\n" % (k, d[k], u.get(k, "")))
            stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
                      for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k and v not in ("", "n/a")}
            fh.write("  warp stalls per issue: " + ", ".join("%s %.2f" % kv for kv in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]) + "\n")


summarise("p3_dist16.ncu-rep", "r2_lu_dist_step_paired_full.txt",
          "# ncu --set full --clock-control none --import-source on -k regex:lu_dist_step_kernel -s 262 -c 2 python bench.py --gpus 1 --steps 1 --warmup 3 --workload swarm --swarm-aircraft 16 --force-partitioned\n"
          "# row-partitioned LU at N = 16,000, world size 1, final kernels of the round: the rank-128 PAIRED visit <0, 2> (56,644 tiles, panels k and k+1 in one visit)\n"
          "# and the THIN launch <0, 1> before it (473 role CTAs: DIAG, U row, L column of panel k+1)")
summarise("p3_residual.ncu-rep", "r2_residual_split_full.txt",
          "# ncu --set full --clock-control none --import-source on -k regex:residual -s 12 -c 2 python bench.py --workload c3 --steps 1 --warmup 3\n"
          "# batches of small scenes (C3: 4,096 cases, N = 180): lean warp-per-row matvec, then the section-lift tail as one thread per row")
lines = [l for l in open(os.path.join(G, "p3_c3_launches.csv")) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
tot = {}
for r in rows:
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("lls::", "").replace("void ", "")
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1
    t[1] += float(r["Metric Value"].replace(",", "")) * 1e-3
allus = sum(v[1] for v in tot.values())
with open(os.path.join(P, "r2_launch_shares_c3.csv"), "w") as fh:
    fh.write("# ncu launch list, about three C3 sweeps (4,096 cases, N = 180, nonlinear): python bench.py --workload c3 --steps 1 --warmup 3\n")
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 (cold-cache, serialised: compare SHARES)\n")
    fh.write("kernel,launches,total_us,avg_us,share\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        fh.write("%s,%d,%.1f,%.1f,%.4f\n" % (k, v[0], v[1], v[1] / v[0], v[1] / allus))
print(open(os.path.join(P, "r2_launch_shares_c3.csv")).read())
