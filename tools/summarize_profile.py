"""Turns the raw artefacts a tools/gpu_profile.sh run leaves in gpurun_out/ into the small, tracked summaries under
profiles/:   python tools/summarize_profile.py <tag> <out_prefix>     e.g.  d  profiles/r01_d
  <out_prefix>_launches.csv   the ncu launch list of one step (kernel, count, total us, share) + every launch
  <out_prefix>_conv.csv       per conv launch: duration, tensor-pipe %, DRAM bytes, DRAM/L2/L1 throughput %
  <out_prefix>_summary.md     both as markdown + the bench line
"""
import collections
import csv
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(REPO, "gpurun_out")
CONV_METRICS = [
    ("gpu__time_duration.sum", "ms"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("dram__bytes_read.sum", "dram_read_GB"),
    ("dram__bytes_write.sum", "dram_write_GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("launch__grid_size", "grid"),
]


def launches(tag):
    rows = list(csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    L = [(re.sub(r"\(.*", "", r[ki]), float(r[vi].replace(",", "")) / 1000.0) for r in rows[hdr + 1:] if len(r) > vi]
    starts = [i for i, (k, _) in enumerate(L) if k.startswith("k_prep_input")]
    step = L[starts[-1]:] if starts else L
    return step


def conv_table(tag, layers):
    rows = list(csv.reader(open(os.path.join(G, "conv_raw_%s.csv" % tag))))
    H, U, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(H)}
    scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}
    out = []
    for r, lay in zip(data, layers):
        d = collections.OrderedDict(op=lay["op"], N=lay["N"], KH=lay["KH"], S=lay["S"], MT=lay.get("MT", ""), resident=lay.get("resident", ""),
                                    bench_ms=lay["ms"], bench_tflops=lay["tflops"])
        for m, name in CONV_METRICS:
            v = r[idx[m]] if m in idx else ""
            if v and (name.endswith("_GB") or name == "ms"):
                v = "%.6f" % (float(v.replace(",", "")) * scale.get(U[idx[m]], 1.0))      # normalise to GB / ms
            d[name] = v
        out.append(d)
    return out


def kernel_rows(path):
    """[(kernel, {metric: value})] in launch order from an ncu --csv launch list with several metrics."""
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ii, ki, mi, ui, vi = H.index("ID"), H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Unit"), H.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    out = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        d = out.setdefault(r[ii], [re.sub(r"\(.*", "", r[ki]), {}])
        d[1][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    return list(out.values())


def cc_summary(tag, prefix, batch=148, match_frames=32):
    """CC stage: per-kernel time and DRAM traffic of the last (warm) label iteration, and of the matching frames."""
    lab = kernel_rows(os.path.join(G, "launches_cc_%s.csv" % tag))
    last = [i for i, (k, _) in enumerate(lab) if k.startswith("k_strip_label")][-1]
    seq = lab[last:last + 3]
    lines = ["kernel,launches,total_us,dram_MB,frames"]
    tot_b = tot_us = 0.0
    for k, m in seq:
        b = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        tot_b += b; tot_us += m["gpu__time_duration.sum"]
        lines.append("%s,1,%.1f,%.2f,%d" % (k, m["gpu__time_duration.sum"], b / 1e6, batch))
    mt = kernel_rows(os.path.join(G, "launches_match_%s.csv" % tag))
    agg = collections.OrderedDict()
    for k, m in mt:
        if k.startswith("k_match"):
            a = agg.setdefault(k, [0, 0.0, 0.0])
            a[0] += 1; a[1] += m["gpu__time_duration.sum"]; a[2] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    for k, (n, us, b) in agg.items():
        lines.append("%s,%d,%.1f,%.2f,%d" % (k, n, us, b / 1e6, n))
    with open(prefix + "_cc_launches.csv", "w") as f:
        f.write("\n".join(lines) + "\n")
        f.write("\n# label rows: python tools/cc_bench.py --batches %d --iters 1 --no-match (third iteration); match rows: --batches %d "
                "--iters 1 (launch counts = frames matched); ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                "dram__bytes_write.sum --clock-control none\n" % (batch, match_frames))
    with open(os.path.join(REPO, "profiles", "cc_traffic.json"), "w") as f:
        json.dump({"dram_bytes_per_frame": tot_b / batch, "label_us_per_frame_under_ncu": tot_us / batch, "batch": batch,
                   "source": os.path.basename(prefix) + "_cc_launches.csv",
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum of k_strip_label + k_resolve + k_crop_fill, per frame"}, f)
    print("wrote", prefix + "_cc_launches.csv")


def main():
    tag, prefix = sys.argv[1], sys.argv[2]
    if os.path.exists(os.path.join(G, "launches_cc_%s.csv" % tag)) and os.path.exists(os.path.join(G, "launches_match_%s.csv" % tag)):
        cc_summary(tag, prefix)
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    step = launches(tag)
    agg = collections.OrderedDict()
    for k, us in step:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(us for _, us in step)
    with open(prefix + "_launches.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct"])
        for k, (n, us) in agg.items():
            w.writerow([k, n, "%.1f" % us, "%.2f" % (100 * us / total)])
        w.writerow([])
        w.writerow(["launch_index", "kernel", "us"])
        for i, (k, us) in enumerate(step):
            w.writerow([i, k, "%.2f" % us])
    layers = json.load(open(os.path.join(G, "layers_%s.json" % tag)))
    conv = conv_table(tag, layers)
    with open(prefix + "_conv.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(conv[0].keys()))
        w.writeheader()
        w.writerows(conv)
    try:
        tot = sum(float(d["dram_read_GB"]) + float(d["dram_write_GB"]) for d in conv) * 1e9
        with open(os.path.join(REPO, "profiles", "conv_traffic.json"), "w") as f:
            json.dump({"dram_bytes_per_step": tot, "launches": len(conv), "source": os.path.basename(prefix) + "_conv.csv",
                       "note": "sum over the conv launches of one 8-frame step of dram__bytes_read.sum + dram__bytes_write.sum (ncu --set full)"}, f)
    except Exception as e:            # units other than Gbyte: leave the file alone
        print("conv_traffic.json not written:", e)
    bench = open(os.path.join(G, "bench_%s.json" % tag)).read().strip()
    with open(prefix + "_summary.md", "w") as f:
        f.write("# profile %s\n\n" % tag)
        f.write("Command: `python tools/profile_step.py` (8 x 1080p frames resident in HBM, 2 steps; second step listed). ncu launch list "
                "(`--metrics gpu__time_duration.sum --clock-control none`): cold-cache, serialised -- compare SHARES.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, us) in agg.items():
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, n, us, 100 * us / total))
        f.write("| **step** | %d | %.1f | |\n\n" % (len(step), total))
        f.write("Per conv launch (`ncu --set full`, same command; bench_* columns are CUDA-event timings from `bench.py --layer-table`):\n\n")
        keys = list(conv[0].keys())
        f.write("| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
        for d in conv:
            f.write("| " + " | ".join(str(d[k])[:9] for k in keys) + " |\n")
        f.write("\nbench.py line of the same build:\n\n```\n%s\n```\n" % bench)
    print("wrote", prefix + "_{launches.csv,conv.csv,summary.md}")


if __name__ == "__main__":
    main()
