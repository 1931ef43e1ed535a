"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_shares.py gpurun_out/launches.csv "<what was run>" > profiles/rNN_launches_step.summary.txt"""
import collections
import csv
import sys


def main(path, note):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        t = tot[r["Kernel Name"][:90]]
        t[0] += us
        t[1] += 1
    total = sum(t[0] for t in tot.values())
    n = sum(t[1] for t in tot.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none, {note}")
    print(f"total device time {total / 1e3:.1f} ms over {n} launches (cold-cache, serialised: compare SHARES)")
    for k, (us, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f"{us / 1e3:10.2f} ms {100 * us / total:6.2f}%  n={c:5d}  avg={us / c:9.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
