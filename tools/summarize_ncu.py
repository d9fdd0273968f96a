"""Turn ncu outputs brought back by gpurun into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/r2_launches.csv profiles/r1_launches_bench.txt
  python tools/summarize_ncu.py full gpurun_out/r2_conv_tc.ncu-rep profiles/r1_conv_tc_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "lts__t_sector_hit_rate.pct"]


def short(k):
    m = re.search(r"\b(k_[a-z0-9_]+)", k)
    if m:
        return "mpb200::" + m.group(1)
    return re.sub(r"<.*", "", k).replace("void ", "")[:60]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u == "s" else v
        rows.append((int(r["ID"]), short(r["Kernel Name"]), v, r["Grid Size"]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, k, v, _g in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if k.startswith("mpb200::"))
    with open(dst, "w") as f:
        f.write(f"# source: {src}\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised:"
                f" compare SHARES, not absolutes)\n# {len(rows)} launches, {tot:.2f} ms total; libmpb200 kernels "
                f"{mine:.2f} ms ({100 * mine / tot:.1f}%)\n\n## share by kernel\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{ms:10.3f} ms {100 * ms / tot:6.2f}%  n={n:5d}  {k}\n")
        f.write("\n## libmpb200 launches in order (id, ms, grid, kernel)\n")
        for i, k, v, g in rows:
            if k.startswith("mpb200::"):
                f.write(f"{i:6d} {v:9.4f} {g:>18s} {k}\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n## {short(d.get('Kernel Name', ''))}  grid={d.get('launch__grid_size')} "
                    f"block={d.get('launch__block_size')}\n")
            for i, h in enumerate(hdr):
                if any(h == k or h.startswith(k + ".") for k in KEYS) and "pct_of_peak_sustained_elapsed" not in h.replace(
                        "avg.pct_of_peak_sustained_elapsed", ""):
                    f.write(f"  {h:75s} {r[i]:>16s} {units[i]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
