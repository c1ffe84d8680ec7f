"""SASS size of every kernel in libthor_b200.so (instruction count x 16 bytes), largest first."""
import os, re, subprocess, sys, tempfile
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "thor_b200", "libthor_b200.so")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", os.path.join(d, cub)], capture_output=True, text=True).stdout
sizes, cur = {}, None
for line in out.split("\n"):
    m = re.match(r"\s*\.section\s+(\S+)", line)
    if m:
        cur = m.group(1) if m.group(1).startswith(".text.") else None
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
        sizes[cur] = sizes.get(cur, 0) + 1
for k, v in sorted(sizes.items(), key=lambda kv: -kv[1])[:12]:
    name = subprocess.run(["c++filt", k[6:].rstrip(",")], capture_output=True, text=True).stdout.strip()
    print("%7d instr %8.1f KB  %s" % (v, v * 16 / 1024, name[:90]))
