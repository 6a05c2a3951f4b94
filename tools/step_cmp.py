import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
rows = []
for k in a.files:
    d = np.abs(a[k].astype(np.float64) - b[k]).max() if a[k].size else 0.0
    rows.append((d, k, a[k].shape))
for d, k, s in sorted(rows, reverse=True)[:25]: print("%.3e %s %s" % (d, k, s))
