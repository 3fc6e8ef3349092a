"""Per-source-line instruction / sample shares of one kernel in an ncu report, in line order.
python tools/ncu_lines_all.py REPORT.ncu-rep KERNEL_REGEX [MIN_PCT]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, agg, hdr = None, collections.OrderedDict(), None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
    a = agg.setdefault((fname, int(r[0]), r[1].strip()), [0.0, 0.0]); a[0] += float(r[i_s] or 0); a[1] += float(r[i_e] or 0)
ts, te = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"kernel {kern}: {te:.0f} warp instructions, {ts:.0f} samples")
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if 100 * a[1] / te >= minp or 100 * a[0] / ts >= minp:
        print(f"{f}:{ln:5d} samples {100*a[0]/ts:5.1f}% inst {100*a[1]/te:5.1f}%  {src[:90]}")
