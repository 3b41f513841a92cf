"""Summaries for profiles/: (1) per-kernel totals of an ncu launch list (gpu__time_duration CSV),
(2) key metrics of each kernel in an ncu --set full report.
    python tools/ncu_summary.py launches gpurun_out/r01_launches.csv
    python tools/ncu_summary.py full gpurun_out/r01_prof_top.ncu-rep
"""
import collections, csv, re, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"hma::", "", name)
        agg[name[:80]][0] += 1
        agg[name[:80]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {sum(v[0] for v in agg.values())} launches, {tot/1e3:.2f} ms of kernel time (cold-cache, serialised by ncu: compare SHARES)")
    print(f"{'kernel':82s} {'n':>5s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:82s} {v[0]:5d} {v[1]:10.1f} {100*v[1]/tot:6.1f}% {v[1]/v[0]:8.1f}")

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "mufu_pipe_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("lts__t_bytes.sum", "l2_bytes"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem")]

def full(path):
    """path: an .ncu-rep (read through `ncu -i`) or the `--page raw --csv` text already exported on the GPU box (a full-set
    report of 10+ launches is larger than gpurun's 64 MiB return limit; its CSV is not)."""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for row in rows[2:]:
        print(re.sub(r"\(.*", "", row[ki])[:100])
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"    {short:20s} {row[i]:>16s} {units[i]}")

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
