"""Per-family summary of an ncu launch list of `python tools/one_step.py B` (developer tool).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum \
        --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_ncu_step.csv python tools/one_step.py 8
    python tools/ncu_summary.py gpurun_out/r2_ncu_step.csv gpurun_out/one_step_families.json profiles/r2_ncu_step_summary.json

The stf_b200 launches of the ncu list are matched, in order and per kernel name, with the family list the same command
wrote (tools/one_step.py); everything else (cuDNN, torch element-wise) is grouped by name.  Output: DRAM bytes, time and
time-weighted tensor-pipe utilisation per family -- what bench.py reports as roofline.traffic / tensor_pipe.ncu."""
import collections
import csv
import json
import re
import sys

KEYS = {"conv_tf32_kernel": "conv_tf32_kernel", "window_attention_tok_kernel": "window_attention_tok_kernel",
        "linear_tf32_kernel": "linear_tf32_kernel", "patch_embed_kernel": "patch_embed_kernel", "swin_mlp_kernel": "swin_mlp_kernel",
        "slice_step_nhwc_kernel": "slice_step_nhwc_kernel", "entropy_bottleneck_kernel": "entropy_bottleneck_kernel",
        "dequantize_kernel": "dequantize_kernel", "window_attention16_kernel": "window_attention_kernel",
        "window_attention64_kernel": "window_attention_kernel"}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    fam = json.load(open(sys.argv[2]))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    idc, kn, mv, mn, mu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        d = launches.setdefault(r[idc], {"name": r[kn]})
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        unit = r[mu]
        if r[mn] == "gpu__time_duration.sum":
            v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v   # -> us
        if r[mn].startswith("dram__bytes"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[mn]] = v
    queues = collections.defaultdict(collections.deque)
    for rec in fam["launches"]:
        base = rec["family"].split("|")[0].split(":")[0]
        queues[base].append(rec)
    out = collections.OrderedDict()
    total_us = 0.0
    for d in launches.values():
        t = d.get("gpu__time_duration.sum", 0.0)
        total_us += t
        key = None
        for pat, base in KEYS.items():
            if pat in d["name"]:
                q = queues.get(base)
                key = q.popleft()["family"] if q else base + "|unmatched"
                break
        if key is None:
            key = "other: " + re.sub(r"[<(].*", "", d["name"])[:60]
        f = out.setdefault(key, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "_tp": 0.0,
                                 "tensor_inst": 0.0})
        f["launches"] += 1
        f["time_us"] += t
        f["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
        f["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
        tp = [v for k, v in d.items() if k.startswith("sm__pipe_tensor_cycles_active")]
        f["_tp"] += (tp[0] if tp else 0.0) * t
        f["tensor_inst"] += d.get("sm__inst_executed_pipe_tensor.sum", 0.0)
    for f in out.values():
        f["tensor_pipe_pct"] = f.pop("_tp") / f["time_us"] if f["time_us"] else None
        f["time_share_pct"] = 100 * f["time_us"] / total_us if total_us else None
        f["dram_gbs"] = (f["dram_read_bytes"] + f["dram_write_bytes"]) / f["time_us"] / 1e3 if f["time_us"] else None
    res = {"batch": fam["batch"], "precision": fam["precision"], "total_kernel_time_us": total_us,
           "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum "
                  "--clock-control none --profile-from-start off, python tools/one_step.py (the second of two eager compress+decompress steps; "
                  "cold-cache, serialised: shares)",
           "families": dict(sorted(out.items(), key=lambda kv: -kv[1]["time_us"]))}
    json.dump(res, open(sys.argv[3], "w"), indent=1)
    for k, f in res["families"].items():
        print(f"{f['time_share_pct']:5.1f}%  {f['time_us'] / 1e3:8.2f} ms  {f['launches']:5d}  dram {f['dram_gbs'] or 0:7.0f} GB/s  "
              f"tensor {f['tensor_pipe_pct'] or 0:5.1f}%  {k}")


if __name__ == "__main__":
    main()
