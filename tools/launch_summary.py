"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) per kernel for profiles/: launches, total and average
device time, share of the captured time.  usage: python tools/launch_summary.py launches.csv "<command that was profiled>" > profiles/x.md
ncu serialises the launches and runs each with a cold cache, so the SHARES are comparable with the bench, not the absolute times."""
import collections
import csv
import re
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        ns = v * {"us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(r[iu], 1.0)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
        name = re.sub(r"cub::(\w+)<.*", r"cub::\1<...>", name)
        tot[name] += ns
        cnt[name] += 1
    T = sum(tot.values())
    print("# ncu launch list: `%s`\n" % cmd)
    print("%d launches, %.1f ms of device time in total (cold-cache, serialised: compare shares).\n" % (sum(cnt.values()), T / 1e6))
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        if v / T < 0.001:
            continue
        print("| `%s` | %d | %.2f | %.1f %% | %.1f |" % (k, cnt[k], v / 1e6, 100 * v / T, v / cnt[k] / 1e3))


if __name__ == "__main__":
    main()
