"""Per-source-line summary of an ncu report: python profiles/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, agg = None, collections.OrderedDict()
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue        # keep the per-line (aggregated) rows only
    i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
    key = (fname, r[0], r[1].strip())
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += float(r[i_s] or 0); a[1] += float(r[i_e] or 0)
ts, te = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"kernel {kern}: {te:.0f} warp instructions, {ts:.0f} samples")
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:>4s} samples {100*a[0]/ts:5.1f}%  inst {100*a[1]/te:5.1f}%  {src[:100]}")
