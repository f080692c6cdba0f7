"""Attribute the warp-stall samples of an ncu --set full capture to CUDA source lines.

python tools/ncu_lines.py <rep> <lib.so> <mangled-kernel-substring> [top]

ncu's csv source page is SASS only; the line table comes from `nvdisasm -g` on the cubin inside the shared library
(built with -lineinfo).  Instruction i of the kernel in the report = instruction i of the same function in the cubin.
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, lib, ksub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
# the library holds one cubin per translation unit: take the one that defines the kernel.  Device functions the kernel
# calls (__noinline__) are separate .text sections; ncu lists them after the kernel's own instructions.
lines = []          # per instruction: (file, line, text)
for cubin in sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    if not any(ln.startswith("//--------------------- .text.") and ksub in ln for ln in dis):
        continue
    sections, name = {}, None
    cur = ("?", 0)
    for ln in dis:
        if ln.startswith("//--------------------- .text."):
            name = ln.split(".text.", 1)[1].split()[0]
            sections[name] = []
            continue
        if name is None:
            continue
        if ln.startswith("//--------------------- "):
            name = None
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            sections[name].append((cur[0], cur[1], m.group(2).strip()))
    kern = [k for k in sections if ksub in k][0]
    lines = list(sections[kern])
    called = set(re.findall(r"\(([A-Za-z0-9_]+)\)", " ".join(t for _, _, t in lines if t.startswith("CALL") or " CALL" in t)))
    extra = [k for k in sections if k != kern and (k in called or os.environ.get("NCU_LINES_ALL_CALLEES"))]
    callee_sections = {k: sections[k] for k in extra}
    break

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
if len(body) != len(lines):
    # try appending the called device functions (every order of the few candidates is cheap to test by length only)
    import itertools
    for r in range(1, len(callee_sections) + 1):
        for combo in itertools.permutations(callee_sections, r):
            if len(lines) + sum(len(callee_sections[k]) for k in combo) == len(body):
                for k in combo:
                    lines += callee_sections[k]
                break
        if len(lines) == len(body):
            break
if len(body) != len(lines):
    print(f"warning: report has {len(body)} SASS lines, cubin function has {len(lines)}", file=sys.stderr)
n = min(len(body), len(lines))
tot = sum(int(r[isamp]) for r in body)
totex = sum(int(r[iex]) for r in body)
agg = {}
for i in range(n):
    key = (lines[i][0], lines[i][1])
    a = agg.setdefault(key, [0, 0, {}, 0])
    a[0] += int(body[i][isamp])
    a[1] += int(body[i][iex])
    a[3] += 1
    for c in stall_cols:
        v = int(body[i][c])
        if v:
            a[2][hdr[c]] = a[2].get(hdr[c], 0) + v
print(f"total samples {tot}, warp instructions executed {totex}, SASS lines {n}")
src_cache = {}
def src(file, line):
    for root in ("ims_toucan_prosody_variance_b200/csrc",):
        p = os.path.join(root, file)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            if 0 < line <= len(src_cache[p]):
                return src_cache[p][line - 1].strip()[:90]
    return ""
print("--- by source line, sorted by samples")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    st = ", ".join(f"{k[6:]} {v}" for k, v in st)
    print(f"{key[0]}:{key[1]:<5d} samples {100 * a[0] / tot:5.2f}%  exec {100 * a[1] / totex:5.2f}%  sass {a[3]:4d}  [{st}]  | {src(*key)}")
print("--- by source line, sorted by executed instructions")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0]}:{key[1]:<5d} exec {100 * a[1] / totex:5.2f}%  samples {100 * a[0] / tot:5.2f}%  sass {a[3]:4d}  | {src(*key)}")
