"""Side-by-side per-layer conv timings of two `bench.py --layer-table` files:  python tools/compare_layers.py a.json b.json"""
import json
import sys

a, b = (json.load(open(p)) for p in sys.argv[1:3])
ta = tb = 0.0
for x, y in zip(a, b):
    print("op %2d/%2d  N %3d KH %d S %2d Sy %d MT %d | %.4f -> %.4f ms  (%+.1f %%)  %7.1f TFLOP/s" %
          (x["op"], y["op"], y["N"], y["KH"], y["S"], y["Sy"], y["MT"], x["ms"], y["ms"], 100 * (y["ms"] / x["ms"] - 1), y["tflops"]))
    ta += x["ms"]
    tb += y["ms"]
print("conv total %.4f -> %.4f ms" % (ta, tb))
