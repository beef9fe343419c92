"""Turn ncu outputs (launch-list csv, .ncu-rep) into the small text summaries committed under profiles/."""
import collections, csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised launches: compare SHARES)  source: {path}",
           f"{'share':>7} {'n':>4} {'avg_us':>9}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{sum(v) / tot * 100:6.2f}% {len(v):4d} {sum(v) / len(v) / 1000:9.2f}  {k}")
    return "\n".join(out)

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]

def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none  source: {path}"]
    for r in rows[2:]:
        out.append(r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                out.append(f"    {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    return "\n".join(out)

if __name__ == "__main__":
    mode, src = sys.argv[1], sys.argv[2]
    print(launches(src) if mode == "launches" else full(src))
