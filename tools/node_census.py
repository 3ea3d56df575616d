"""Per-kernel-family census of an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of tools/profile_step.py:
launches, summed time and share, separately for the B = 2 evaluation and the B = 2*forward_batch chunk.
Usage: python tools/node_census.py profiles/r02_launches_v45_B2_B100.csv > profiles/r02_node_census_v45.md"""
import collections
import csv
import re
import sys


def family(name):
    m = re.search(r"(?:aedit::)?(?:<unnamed>::|\(anonymous namespace\)::)?(\w+_kernel)\s*[<(]", name)
    return m.group(1) if m else "torch / other"


rows = []
with open(sys.argv[1]) as f:
    for x in csv.DictReader(l for l in f if not l.startswith("==")):
        if x.get("Metric Name") == "gpu__time_duration.sum":
            g = tuple(int(v) for v in re.findall(r"\d+", x["Grid Size"]))
            rows.append((x["Kernel Name"], g, float(x["Metric Value"].replace(",", ""))))
big = next(i for i, r in enumerate(rows) if "gemm" in r[0] and r[1][0] * r[1][1] * r[1][2] > 1500)
start = max(i for i in range(big) if "timestep_embedding" in rows[i][0])     # first kernel of the large-batch evaluation
print(f"# Kernel census of one evaluation per batch size ({sys.argv[1]})\n")
print("ncu serialises the launches and runs each with cold caches: compare shares and counts, not absolute times "
      "(inside the captured graph the B = 2 evaluation takes 5.7-5.9 ms, the B = 100 chunk 78.8 ms).")
for label, sub in (("B = 2 evaluation (reverse step)", rows[:start]), ("large-batch forward chunk", rows[start:])):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, _, ns in sub:
        a = agg[family(name)]
        a[0] += 1
        a[1] += ns
    tot, n = sum(v[1] for v in agg.values()), sum(v[0] for v in agg.values())
    print(f"\n## {label}: {n} launches, {tot / 1e6:.2f} ms summed\n")
    print("| kernel | launches | total µs | share | mean µs |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.0f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0] / 1e3:.1f} |")
