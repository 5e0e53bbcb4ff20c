"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family (developer tool).
   python tools/launch_summary.py gpurun_out/launches.csv [top]"""
import collections
import csv
import re
import sys


def family(name):
    if "stf::" in name:
        m = re.search(r"(\w+_kernel)", name)
        return "stf_b200 " + (m.group(1) if m else name[:40])
    if "cutlass" in name or "cudnn" in name or "convolve" in name or "nhwc" in name.lower() or "conv" in name.lower():
        return "cuDNN convolution (" + re.sub(r"[<(].*", "", name)[:48] + ")"
    if "at::" in name:
        if "gpu_kernel_impl_nocast" in name and "vectorized" not in name:
            return "torch strided element-wise (copies / broadcast adds)"
        for key, label in (("Gelu", "torch GELU"), ("CatArray", "torch cat"), ("tanh", "torch tanh"), ("layer_norm", "torch LayerNorm"),
                           ("m_kernel", "torch LayerNorm")):
            if key in name:
                return label
        return "torch other element-wise"
    return re.sub(r"[<(].*", "", name)[:60]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, mv, mn, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
        f = family(r[kn])
        agg[f][0] += 1
        agg[f][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold cache: shares only)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] / tot * 100:5.1f}%  {v[1] / 1e3:8.2f} ms  {v[0]:5d} launches  {k}")


if __name__ == "__main__":
    main()
