"""Dev tool: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> shares of the step."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    n = re.sub(r"\(anonymous namespace\)::", "", row["Kernel Name"])
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n).replace("at::native::", "")[:72]
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if not k.startswith(("at::", "nvjet", "cutlass", "cublas")))
n_ours = sum(v[0] for k, v in agg.items() if not k.startswith(("at::", "nvjet", "cutlass", "cublas")))
n_all = sum(v[0] for v in agg.values())
print(f"{n_all} launches, {tot:.0f} us of kernel time; this repo's kernels: {n_ours} launches, {ours:.0f} us ({100 * ours / tot:.1f} %)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 50]:
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}% {v[0]:5d}  {k}")
