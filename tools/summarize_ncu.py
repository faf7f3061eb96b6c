#!/usr/bin/env python
"""Turn `ncu --page raw --csv` exports of tools/profile_stages.py (two launches per stage variant,
the second one is kept) into the small summaries under profiles/ that bench.py reads for
`roofline.traffic`.  Usage: summarize_ncu.py <raw.csv> <out.json> <cfg3|cfg4|onepass_cfg3|...> [points per launch]"""
import csv
import json
import sys

NAMES = {"cfg3": ["F2d", "B2d[G]", "B2d[I]", "BB2d[GO]", "BB2d[IO]", "BBB2d[IO+X2]"],
         "cfg4": ["F3d", "B3d[G]", "B3d[I]", "BB3d[GO]", "BB3d[IO]", "BBB3d[IO+X2]"],
         # tools/profile_fused.py: the three kernels of the fused jet step
         "fused_cfg3": ["JET2d[fwd]", "HEAD2d", "JET2d[bwd]"],
         "fused_cfg4": ["JET3d[fwd]", "HEAD3d", "JET3d[bwd]"],
         # tools/profile_onepass.py: the one-pass kernel, 2^22 binned points per launch
         "onepass_cfg3": ["ONEPASS2d"], "onepass_cfg4": ["ONEPASS3d"]}
POINTS = {"onepass_cfg3": 2 ** 22, "onepass_cfg4": 2 ** 22}
NOTE = {"cfg3": "cells [4,16,256,256], 2^20 points, cosine multicell",
        "cfg4": "cells [4,16,64,64,64], 2^22 points, smoothstep multicell",
        "fused_cfg3": "tools/profile_fused.py: cells [4,16,256,256], 2^20 points, cosine multicell",
        "fused_cfg4": "tools/profile_fused.py: cells [4,16,64,64,64], 2^22 points, smoothstep multicell",
        "onepass_cfg3": "tools/profile_onepass.py: cells [4,16,256,256], 2^22 binned points, cosine multicell",
        "onepass_cfg4": "tools/profile_onepass.py: cells [4,16,64,64,64], 2^22 binned points, smoothstep multicell"}
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main(raw, out, cfg, points=None):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def val(r, k):
        i = hdr.index(k)
        return float(data[r][i].replace(",", "")) * SCALE.get(units[i], 1)

    summary = {}
    for r in range(1, len(data), 2):
        summary[NAMES[cfg][r // 2]] = {
            "ncu_duration_us": val(r, "gpu__time_duration.sum"),
            "dram_bytes_read": val(r, "dram__bytes_read.sum"),
            "dram_bytes_write": val(r, "dram__bytes_write.sum"),
            "traffic_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
            "l2_to_l1_read_bytes": val(r, "l1tex__m_xbar2l1tex_read_bytes.sum"),
            "l1_to_l2_write_bytes": val(r, "l1tex__m_l1tex2xbar_write_bytes.sum"),
            "registers": val(r, "launch__registers_per_thread"),
            "l1tex_throughput_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
            "l1tex2xbar_req_cycles_pct": val(r, "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "lts_throughput_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": val(r, "smsp__inst_executed.sum"),
        }
    doc = {"source": "ncu --set full --clock-control none, %s (%s); second launch of each kernel" % (cfg, NOTE[cfg]),
           "kernels": summary}
    if cfg in POINTS:
        doc["points_per_launch"] = int(points) if points else POINTS[cfg]
        if points:
            doc["source"] = doc["source"].replace("2^22 binned points", "%d binned points" % int(points))
    json.dump(doc, open(out, "w"), indent=1)
    for k, v in summary.items():
        print("%-14s %8.1f us  dram %7.1f MB  L2->L1 %7.1f MB  l1tex %4.1f%% req %4.1f%% lts %4.1f%% dram %4.1f%% issue %4.1f%% regs %d"
              % (k, v["ncu_duration_us"], v["traffic_bytes"] / 1e6, v["l2_to_l1_read_bytes"] / 1e6, v["l1tex_throughput_pct"],
                 v["l1tex2xbar_req_cycles_pct"], v["lts_throughput_pct"], v["dram_throughput_pct"], v["issue_active_pct"], v["registers"]))


if __name__ == "__main__":
    main(*sys.argv[1:5])
