"""ncu launch list (gpu__time_duration + dram bytes, --csv) -> per-kernel summary + the `traffic` figure bench.py reports.

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r02 [source_hash]
writes <prefix>_launches_summary.csv, <prefix>_launches_full.csv (the raw list) and <prefix>_traffic.json, stamped with
the hash of the CUDA sources the capture was taken from (bench.py flags the figure as stale when the sources change).
"""
import collections
import csv
import json
import re
import shutil
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("cvvae::", "")
    return name[:70]


def main(src, prefix, src_hash=None):
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"])})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"].startswith("gpu__time_duration"):
            d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3}.get(unit, 1e-6)
        else:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["rd" if "read" in r["Metric Name"] else "wr"] = v * scale
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["kernel"], {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["ms"] += d.get("ms", 0.0)
        a["rd"] += d.get("rd", 0.0)
        a["wr"] += d.get("wr", 0.0)
    total = sum(a["ms"] for a in agg.values())
    with open(prefix + "_launches_summary.csv", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv\n"
                f"#   command: python bench.py --steps 1 --warmup 3 --no-cpu-baseline   ({len(per)} launches, serialised, cold-cache, not power-capped:\n"
                f"#   compare SHARES with bench.py roofline.share_of_step)\n# total {total:.2f} ms\n")
        f.write("kernel,launches,time_ms,share,dram_read_GB,dram_write_GB,GB_per_launch\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            f.write(f"\"{k}\",{a['n']},{a['ms']:.2f},{a['ms'] / total:.4f},{a['rd'] / 1e9:.2f},{a['wr'] / 1e9:.2f},"
                    f"{(a['rd'] + a['wr']) / a['n'] / 1e9:.3f}\n")
    shutil.copyfile(src, prefix + "_launches_full.csv")
    conv = [(k, a) for k, a in agg.items() if k.startswith("conv_tc") or k.startswith("conv_stk")]
    if src_hash is None:
        sys.path.insert(0, ".")
        import bench
        src_hash = bench.source_hash()
    n = sum(a["n"] for _, a in conv)
    tot = sum(a["rd"] + a["wr"] for _, a in conv)
    with open(prefix + "_traffic.json", "w") as f:
        json.dump({"source": f"{prefix}_launches_summary.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum over "
                             f"{len(per)} steady-state launches of bench.py)",
                   "kernel": " + ".join(sorted({k.split('<')[0] for k, _ in conv})), "launches": n, "dram_bytes_total": tot,
                   "dram_bytes_per_launch": tot / max(n, 1), "source_hash": src_hash}, f, indent=1)
    print(open(prefix + "_launches_summary.csv").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
