"""Static SASS instruction count per source line of one kernel: python tools/sass_lines.py MANGLED_NAME [TOP]"""
import re, collections, subprocess, sys, tempfile, os, glob
name = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "shoulder_b200", "libshoulder_b200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [c for c in glob.glob(d + "/*.cubin") if os.path.basename(c).startswith("shb_kernels.")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
cur, cnt, on = None, collections.Counter(), False
for l in dis.split("\n"):
    if l.startswith(".text."): on = name in l
    if not on: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+[A-Z@]", l): cnt[cur] += 1
print("total", sum(cnt.values()), "instructions =", sum(cnt.values()) * 16 / 1024, "KB")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]: print(k, v)
