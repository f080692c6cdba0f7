"""Summarise ncu artefacts into small text files for profiles/ (run in the authoring container).

    python tools/ncu_summary.py full  <report.ncu-rep> <out.txt>     # one `--set full` capture
    python tools/ncu_summary.py list  <launches.csv>  <out.txt>      # `--metrics gpu__time_duration.sum` launch list
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
]
STALLS = "smsp__average_warps_issue_stalled_"


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {rep}"]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        lines.append(f"\n## kernel {rec.get('Kernel Name', '?')[:100]}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
        for k in KEYS:
            if k in rec:
                lines.append(f"{k:90s} {rec[k]:>18s} {units[hdr.index(k)]}")
        stalls = sorted(((float(rec[h]), h) for h in hdr if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")
                         and rec[h] not in ("", "n/a")), reverse=True)
        lines.append("top warp stall reasons (warps stalled per issue-active cycle):")
        for v, h in stalls[:6]:
            lines.append(f"  {h[len(STALLS):-len('_per_issue_active.ratio')]:30s} {v:.3f}")
    open(out, "w").write("\n".join(lines) + "\n")


def launch_list(path, out):
    agg = collections.OrderedDict()
    total = 0.0
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    per = []
    for r in rows[1:]:
        ns = float(r[vi].replace(",", ""))
        name = r[ki].split("(")[0].replace("void ", "")
        per.append((name, r[gi], ns))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    lines = [f"# ncu launch list {path}: {len(per)} launches, {total / 1e6:.3f} ms total (cold-cache, serialised: compare shares)",
             f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}"]
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name[:70]:70s} {n:8d} {ns / 1e6:10.3f} {100 * ns / total:6.1f}%")
    lines.append("\n# per launch (order of execution)")
    for name, grid, ns in per:
        lines.append(f"{name[:60]:60s} grid {grid:>16s} {ns / 1e3:10.1f} us")
    open(out, "w").write("\n".join(lines) + "\n")


def step_list(path, out):
    """Launch list with time + DRAM bytes per launch (three metrics per kernel id); also prints the mean traffic."""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ii, ki, bi, mi, vi = (hdr.index(k) for k in ("ID", "Kernel Name", "Block Size", "Metric Name", "Metric Value"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "").replace("tb200::", ""), "block": r[bi]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    tot_ns = sum(d["gpu__time_duration.sum"] for d in per.values())
    rd = sum(d["dram__bytes_read.sum"] for d in per.values())
    wr = sum(d["dram__bytes_write.sum"] for d in per.values())
    lines = [f"# ncu launch list of ONE timed step of bench.py (tools/gpu_final.sh: -k regex of the conv kernels, -s warm-up launches, -c launches per step): {len(per)} launches",
             "# metrics: gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum (cold-cache, serialised: compare shares)",
             f"# total {tot_ns / 1e6:.3f} ms, DRAM read {rd / 1e9:.2f} GB, write {wr / 1e9:.2f} GB -> {(rd + wr) / len(per) / 1e6:.1f} MB per launch",
             f"{'#':>3s} {'kernel':44s} {'block':>13s} {'us':>10s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s}"]
    for i, d in enumerate(per.values()):
        ns = d["gpu__time_duration.sum"]
        lines.append(f"{i:3d} {d['name'][:44]:44s} {d['block']:>13s} {ns / 1e3:10.1f} {100 * ns / tot_ns:5.1f}% "
                     f"{d['dram__bytes_read.sum'] / 1e6:9.1f} {d['dram__bytes_write.sum'] / 1e6:9.1f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"mean DRAM traffic per launch: {int((rd + wr) / len(per))} bytes")


if __name__ == "__main__":
    {"full": full, "list": launch_list, "steplist": step_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
