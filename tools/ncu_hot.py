"""Hot SASS regions of an ncu --set full capture: python tools/ncu_hot.py <rep> [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[isamp]) for r in body); totex = sum(int(r[iex]) for r in body)
print(f"total samples {tot}, warp instructions executed {totex}, SASS lines {len(body)}")
# aggregate in windows of 64 instructions
W = 64
for w0 in range(0, len(body), W):
    win = body[w0:w0 + W]
    s = sum(int(r[isamp]) for r in win); ex = sum(int(r[iex]) for r in win)
    if s < tot * 0.01: continue
    st = {hdr[i]: sum(int(r[i]) for r in win) for i in stall_cols}
    top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    ops = {}
    for r in win:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
        ops[op] = ops.get(op, 0) + 1
    topops = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{w0:5d}] samples {100*s/tot:5.1f}%  exec {100*ex/totex:5.1f}%  stalls {top3}  ops {topops}")
print("--- hottest single instructions")
for r in sorted(body, key=lambda r: -int(r[isamp]))[:top]:
    idx = body.index(r)
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx:6d} {100*int(r[isamp])/tot:5.2f}% {r[isrc][:70]:70s} {st}")
