#!/usr/bin/env python
"""Summarise an `ncu --set full` report: per-kernel medians of the metrics the rooflines quote.

    python tools/ncu_summary.py gpurun_out/full_c2.ncu-rep profiles/r1_ncu_full_summary.csv [profiles/traffic.json]

traffic.json maps bench.py's kernel names to DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum, median over the captured launches).
"""
import csv
import io
import json
import statistics
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum"]
NAMES = {"nms_kernel": "detect", "mnn_tc_kernel<(int)3": "mnn_similarity_fp16x3", "mnn_tc_kernel<3": "mnn_similarity_fp16x3",
         "voxel_scatter_kernel": "voxel_scatter", "voxel_tile_kernel": "voxel_tile", "detect_kernel": "detect",
         "sample_bilinear_slab_kernel": "sample", "sample_kernel": "sample", "mnn_tc_kernel<(int)1": "mnn_similarity_tf32x3",
         "mnn_tc_kernel<1": "mnn_similarity_tf32x3", "mnn_tc_kernel<(int)0": "mnn_similarity_bf16",
         "mnn_tc_kernel<0": "mnn_similarity_bf16", "mnn_tc_kernel<(int)2": "mnn_similarity_fp16x3_inkernel",
         "mnn_tc_kernel<2": "mnn_similarity_fp16x3_inkernel", "mnn_fp32_kernel": "mnn_similarity_fp32"}


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        per.setdefault(r[col["Kernel Name"]], []).append(r)
    have = [m for m in METRICS if m in col]
    traffic = {}
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches"] + [f"{m} [{units[col[m]]}]" for m in have])
        for k, rs in per.items():
            med = {m: statistics.median(float(r[col[m]].replace(",", "") or 0) for r in rs) for m in have}
            w.writerow([k, len(rs)] + [f"{med[m]:.6g}" for m in have])
            for frag, name in NAMES.items():
                if frag in k:
                    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
                    b = sum(med[m] * scale.get(units[col[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    traffic[name] = int(b)
    if len(sys.argv) > 3:
        old = {}
        try:
            old = json.load(open(sys.argv[3]))
        except Exception:
            pass
        old.update(traffic)
        json.dump(old, open(sys.argv[3], "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
