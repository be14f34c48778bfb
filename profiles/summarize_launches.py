"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
usage: python profiles/summarize_launches.py launches.csv "title" "command" > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [l for l in open(path, errors="replace") if l.startswith('"')]
    rd = csv.DictReader(rows)
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        m = re.search(r"pf_kernel<pv::(\w+)", name)
        name = m.group(1) if m else re.sub(r"\(.*", "", name).replace("pv::", "")
        name = re.sub(r"<.*", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        agg[name][0] += 1; agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# {title}\n\nCommand (on the B200 box): `{cmd}`\n\n{n} launches captured, {tot:.3f} ms of kernel time. Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")


if __name__ == "__main__":
    main()
