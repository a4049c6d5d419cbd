"""Turn an `ncu --csv --metrics ...` log into a per-launch table (stdout, markdown) and, with --gemm-traffic, into
profiles/ncu_gemm_traffic.json (read by bench.py for roofline.traffic).

    python tools/ncu_to_json.py gpurun_out/r2c3/ncu_hbm.csv                      # table of every captured launch
    python tools/ncu_to_json.py gpurun_out/r2c3/ncu_gemm.csv --gemm-traffic M N K  # + JSON for bench.py
"""
import collections, csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(rows)
    hdr = next(rd)
    idx = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rd:
        try:
            key = r[idx["ID"]]
            d = launches.setdefault(key, {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", ""),
                                          "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]})
            d[r[idx["Metric Name"]]] = (float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]])
        except Exception:
            continue
    return list(launches.values())


def to_bytes(v):
    val, unit = v
    return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v):
    val, unit = v
    return val * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit, 1e-3)


def main():
    path = sys.argv[1]
    L = parse(path)
    print(f"| # | kernel | grid | time us | DRAM read MB | DRAM write MB | DRAM GB/s | dram_throughput % of peak |")
    print("|---|---|---|---|---|---|---|---|")
    for i, d in enumerate(L):
        t = to_us(d["gpu__time_duration.sum"]) if "gpu__time_duration.sum" in d else float("nan")
        rd = to_bytes(d["dram__bytes_read.sum"]) if "dram__bytes_read.sum" in d else float("nan")
        wr = to_bytes(d["dram__bytes_write.sum"]) if "dram__bytes_write.sum" in d else float("nan")
        pct = next((v[0] for k, v in d.items() if k.startswith("dram__throughput") or k.startswith("gpu__dram_throughput")), float("nan"))
        print(f"| {i} | `{d['kernel'][:60]}` | {d['grid']} | {t:.1f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | {(rd + wr) / t / 1e3:.0f} | {pct:.1f} |")
    if "--gemm-traffic" in sys.argv:
        k = sys.argv.index("--gemm-traffic")
        M, N, K = (int(x) for x in sys.argv[k + 1:k + 4])
        g = [d for d in L if "gemm_bf16_kernel" in d["kernel"]][-1]
        rd, wr = to_bytes(g["dram__bytes_read.sum"]), to_bytes(g["dram__bytes_write.sum"])
        out = {"shape": [M, N, K], "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
               "time_us": round(to_us(g["gpu__time_duration.sum"]), 2), "algorithmic_operand_bytes": 2 * (M * K + N * K), "output_bytes": 2 * M * N,
               "note": f"dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of gemm_bf16_kernel M{M} N{N} K{K} (ncu --clock-control none; "
                       f"{os.path.relpath(path, ROOT)}); algorithmic operand bytes {2 * (M * K + N * K) / 1e6:.1f} MB + {2 * M * N / 1e6:.1f} MB output "
                       "(the output mostly stays in the 126 MB L2 at kernel end)"}
        json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_gemm_traffic.json"), "w"), indent=1)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
