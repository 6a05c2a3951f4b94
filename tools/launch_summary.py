"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log into the per-kernel table kept under profiles/.
usage: python tools/launch_summary.py gpurun_out/launches.csv "title / command text" > profiles/rNN_launches.md"""
import collections, csv, sys
path, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in rows:
    k = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"])
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
print("# Launch list (ncu gpu__time_duration.sum per launch, aggregated by kernel)\n")
print(note + "\n")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
    print("| `%s` | %d | %.1f | %.2f | %.1f %% |" % (k[:72], n, t / 1e3, t / n / 1e3, 100 * t / tot))
g = sum(t for k, (n, t) in agg.items() if "gemm_t" in k)
print("\nTotal %.1f us over %d launches; `gemm_tma_kernel` + `gemm_tc_kernel` family: %.1f %% of device time." % (tot / 1e3, len(rows), 100 * g / tot))
