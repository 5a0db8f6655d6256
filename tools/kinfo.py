"""Compile one .cu of magma_b200/csrc for sm_100a and print, per kernel: registers, spills, stack, static SASS
instruction count (and optionally an opcode histogram).   python tools/kinfo.py lu_small_sq.cu [substr] [--hist]"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "magma_b200", "csrc")
src = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
obj = "/tmp/kinfo_" + os.path.basename(src) + ".o"
r = subprocess.run(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v",
                    "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj],
                   capture_output=True, text=True)
if r.returncode: print(r.stderr); sys.exit(1)
info = {}
cur = None
for line in r.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m: cur = m.group(1); info[cur] = {}
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur: info[cur].update(stack=int(m.group(1)), sst=int(m.group(2)), sld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and cur: info[cur]["regs"] = int(m.group(1))
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cnt = collections.Counter(); hist = collections.defaultdict(collections.Counter); fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if m and fn:
        cnt[fn] += 1
        t = m.group(1).split()
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        hist[fn][".".join(op.split(".")[:2])] += 1
for k, v in info.items():
    if sub and sub not in k: continue
    dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("void mb200::(anonymous namespace)::", "").split("(")[0]
    print(f"{dem:60s} regs {v.get('regs'):4d} spill {v.get('sst',0):5d}/{v.get('sld',0):5d} stack {v.get('stack',0):4d} sass {cnt.get(k,0):6d}")
    if "--hist" in sys.argv:
        print("    " + "  ".join(f"{o}:{c}" for o, c in hist[k].most_common(28)))
