"""Turns the outputs of tools/profile_round.sh into the tracked artefacts under profiles/:
   python tools/ncu_summary.py TAG      (reads gpurun_out/TAG_*, writes profiles/TAG_*)"""
import csv, json, shutil, subprocess, sys, collections
from pathlib import Path
tag = sys.argv[1]
root = Path(__file__).resolve().parent.parent
src, dst = root / "gpurun_out", root / "profiles"
shutil.copy(src / f"{tag}_bench.json", dst / f"{tag}_bench.json")
shutil.copy(src / f"{tag}_launches.csv", dst / f"{tag}_launches.csv")
# ---- launch list: per-kernel totals and shares
tot, cnt = collections.Counter(), collections.Counter()
big = collections.defaultdict(list)
for r in csv.reader(open(src / f"{tag}_launches.csv")):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        name = r[4].split("(")[0]
        tot[name] += float(r[14]) / 1e3; cnt[name] += 1; big[name].append(float(r[14]) / 1e3)
s = sum(tot.values())
with open(dst / f"{tag}_launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400  python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-parity\n")
    f.write("(cold-cache serialised times: compare shares)\n\n")
    for k, v in tot.most_common():
        f.write(f"{k:70s} launches={cnt[k]:4d} total_us={v:10.1f} share={100 * v / s:5.1f}%  full_size_launch_us={max(big[k]):8.1f}\n")
    f.write("\n(full_size_launch_us: the longest launch of each kernel = the device-resident leg's batch of 32 bones; the e2e leg\n"
            " runs the same kernels on 8 groups of 4 bones)\n")
    cfg3 = src / f"{tag}_cfg3_launches.csv"
    if cfg3.exists():
        b3 = collections.defaultdict(list)
        for r in csv.reader(open(cfg3)):
            if len(r) > 14 and r[12] == "gpu__time_duration.sum":
                b3[r[4].split("(")[0]].append(float(r[14]) / 1e3)
        f.write("\n--workload cfg3 (519,040 triangles, 8,192 planes), longest launch per kernel, us:\n")
        for k, v in sorted(b3.items(), key=lambda kv: -max(kv[1])):
            f.write(f"{k:70s} {max(v):8.1f}\n")
# ---- full capture: key metrics per kernel + traffic json
rep = str(src / f"{tag}_full.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]
keys = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
keys += [k for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
units = rows[1]
traffic = {"command": "python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-parity", "workload": "cfg2, 32 bones x 2048 planes, N=360, A=360", "kernels": {}}
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_s(v, u):
    return float(v) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1e-9)
with open(dst / f"{tag}_ncu_full_summary.txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:'k_stitch_group|k_stitch_list|k_resample|k_intersect' -s 12 -c 6  python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-parity\n")
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        f.write("----\n  Kernel Name  " + name + "\n")
        for k in keys:
            if k in h:
                v = r[h.index(k)]
                try:
                    if "stalled" in k and float(v) < 0.2: continue
                except ValueError: pass
                f.write(f"  {k:86s}{v} {units[h.index(k)]}\n")
        rd = to_bytes(r[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_read.sum")])
        wr = to_bytes(r[h.index("dram__bytes_write.sum")], units[h.index("dram__bytes_write.sum")])
        short = name.replace("void ", "").split("(")[0]
        rec = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
               "duration_s_under_ncu": to_s(r[h.index("gpu__time_duration.sum")], units[h.index("gpu__time_duration.sum")])}
        if short not in traffic["kernels"] or rec["duration_s_under_ncu"] > traffic["kernels"][short]["duration_s_under_ncu"]:
            traffic["kernels"][short] = rec          # several launches of a kernel per step (head / bulk of the stitch order): keep the large one
json.dump(traffic, open(dst / f"{tag}_traffic.json", "w"), indent=1)
# ---- per-line hot spots
for kern in ("k_resample", "k_stitch_group", "k_stitch_list", "k_intersect"):
    o = subprocess.run([sys.executable, str(root / "profiles" / "ncu_lines.py"), rep, kern, "25"], capture_output=True, text=True).stdout
    open(dst / f"{tag}_lines_{kern}.txt", "w").write(o)
print("wrote profiles/" + tag + "_*")
