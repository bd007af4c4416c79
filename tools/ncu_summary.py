"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list per kernel."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, vi, idi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Metric Unit")
    data = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = data.setdefault(r[idi], {"k": r[ki]})
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d[r[mi]] = v * scale          # time in us, bytes in bytes
    return data


def short(k):
    return k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]


def main():
    data = load(sys.argv[1])
    if "--step" in sys.argv:                 # drop model construction / optimizer set-up: keep launches from the first pack_params on
        sys.argv.remove("--step")
        keys = list(data)
        first = next(i for i, k in enumerate(keys) if "pack_params" in data[k]["k"])
        data = collections.OrderedDict((k, data[k]) for k in keys[first:])
    if "--gemm-traffic" in sys.argv:         # per-launch DRAM traffic of the tcgen05 GEMM family -> JSON read by bench.py (roofline.traffic)
        import json
        i = sys.argv.index("--gemm-traffic")
        out, src = sys.argv[i + 1], sys.argv[i + 2]
        del sys.argv[i:i + 3]
        g = [d for d in data.values() if "gemm_tc_kernel" in d["k"]]
        json.dump({"dram_bytes_per_launch": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in g) / len(g),
                   "launches": len(g), "avg_us_cold_serialised": sum(d.get("gpu__time_duration.sum", 0) for d in g) / len(g), "source": src},
                  open(out, "w"), indent=1)
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for d in data.values():
        a = agg[short(d["k"])]
        a[0] += d.get("gpu__time_duration.sum", 0)
        a[1] += 1
        a[2] += d.get("dram__bytes_read.sum", 0)
        a[3] += d.get("dram__bytes_write.sum", 0)
    tot = sum(a[0] for a in agg.values())
    print(f"{len(data)} launches, total {tot / 1e3:.2f} ms (cold-cache, serialised: compare shares)")
    print(f"{'ms':>9} {'share':>6} {'n':>5} {'rd GB':>8} {'wr GB':>8} {'GB/s':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        bw = (a[2] + a[3]) / a[0] / 1e3 if a[0] else 0
        print(f"{a[0] / 1e3:9.3f} {100 * a[0] / tot:5.1f}% {a[1]:5d} {a[2] / 1e9:8.3f} {a[3] / 1e9:8.3f} {bw:7.0f}  {k}")
    if len(sys.argv) > 3:
        pat = sys.argv[3]
        sel = sorted((d for d in data.values() if pat in d["k"]), key=lambda d: -d.get("gpu__time_duration.sum", 0))
        for d in sel[:8]:
            t = d.get("gpu__time_duration.sum", 0)
            rd, wr = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
            print(f"  {t:9.1f} us rd {rd / 1e6:8.1f} MB wr {wr / 1e6:8.1f} MB {(rd + wr) / t / 1e3 if t else 0:7.0f} GB/s  {short(d['k'])}")


if __name__ == "__main__":
    main()
