"""Turns the raw ncu outputs of scripts/gpu_j.sh (gpurun_out/j_*) into the tracked summaries under profiles/.
usage: python scripts/summarise_profiles.py [prefix=j] [round_tag=r1]"""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
pre = sys.argv[1] if len(sys.argv) > 1 else "j"
tag = sys.argv[2] if len(sys.argv) > 2 else "r1"

# ---- launch list ------------------------------------------------------------------------------------------------------
lines = [l for l in open(os.path.join(G, pre + "_launches.csv")) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("lls::", "").replace("void ", "")
with open(os.path.join(P, tag + "_launches_c2.csv"), "w") as fh:
    fh.write("id,kernel,block,grid,duration_ns\n")
    for i, r in enumerate(rows):
        fh.write('%d,"%s","%s","%s",%d\n' % (i, short(r["Kernel Name"]), r["Block Size"], r["Grid Size"], int(float(r["Metric Value"].replace(",", "")))))
tot = {}
for r in rows:
    k = short(r["Kernel Name"])
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1
    t[1] += float(r["Metric Value"].replace(",", "")) * 1e-3
allus = sum(v[1] for v in tot.values())
with open(os.path.join(P, tag + "_launch_shares_c2.csv"), "w") as fh:
    fh.write("# ncu launch list, one C2 solve (N=4000): python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n")
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1100 (r1: -s 3500 -c 900; about 1.3 solves; cold-cache, serialised: compare SHARES)\n")
    fh.write("kernel,launches,total_us,share\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        fh.write("%s,%d,%.1f,%.4f\n" % (k, v[0], v[1], v[1] / allus))

# ---- full captures ------------------------------------------------------------------------------------------------------
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
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            fh.write("kernel: %s  grid %s block %s\n" % (d["Kernel Name"][:100], d["Grid Size"], d["Block Size"]))
            for k in KEYS:
                if k in d:
                    fh.write("  %-75s %s %s\n" % (k, d[k], u.get(k, "")))
            stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
                      for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k and v not in ("", "n/a")}
            fh.write("  warp stalls per issue: " + ", ".join("%s %.2f" % kv for kv in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]) + "\n")
summarise(pre + "_lu.ncu-rep", tag + "_lu_step_full.txt",
          "# ncu --set full --clock-control none --import-source on -k regex:lu_step_kernel -s 198 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline  (C2, N=4000; block step 8 of the 4th factorisation)")
summarise(pre + "_influence.ncu-rep", tag + "_influence_full.txt",
          "# ncu --set full --clock-control none --import-source on -k regex:influence_kernel -s 4 -c 2 (r1: -s 2 -c 1) python bench.py --steps 1 --warmup 3 --no-cpu-baseline  (C2, N=4000; r2: <2> = plain tiles, <1> = blend tiles of the same grid)")
if os.path.exists(os.path.join(G, pre + "_dist.ncu-rep")):
    summarise(pre + "_dist.ncu-rep", tag + "_lu_dist_step_full.txt",
              "# ncu --set full --clock-control none --import-source on -k regex:lu_dist_step_kernel -s 72 -c 1 python scripts/dist_check.py c2_traditional_n4000  (row-partitioned LU, world size 1, N=4000; a bulk block step of 2916 tiles: compare with the lu_step_kernel launch of the same size)")
if os.path.exists(os.path.join(G, pre + "_small.ncu-rep")):
    summarise(pre + "_small.ncu-rep", tag + "_lu_small_case_full.txt",
              "# ncu --set full --clock-control none --import-source on -k regex:lu_small_case_kernel -s 30 -c 1 python bench.py --workload c3 --steps 1 --warmup 3  (4,096 cases of 3 block rows, one CTA per case)")
print(open(os.path.join(P, tag + "_launch_shares_c2.csv")).read())
print(open(os.path.join(P, tag + "_lu_step_full.txt")).read())
print(open(os.path.join(P, tag + "_influence_full.txt")).read())
