"""Kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (markdown table on stdout)."""
import collections
import csv
import sys


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hdr]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0]
        t = float(r[mv].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        n += 1
    total = sum(v[1] for v in agg.values())
    print("# %s\n" % title)
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, v[0], v[1], 100.0 * v[1] / total))
    print("\n%d launches, %.1f us in total (cold-cache and serialised under ncu: compare shares, not absolutes)." % (n, total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
