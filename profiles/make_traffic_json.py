"""Regenerates profiles/dominant_kernel_traffic.json (read by bench.py for roofline.traffic / l2_frac) from an
`ncu --set full` capture of the bench-setting apply and the L2 micro-benchmark.

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/make_traffic_json.py /tmp/raw.csv gpurun_out/<l2_bandwidth output>.json [ref_level] [solve_path]

Per kernel class (bench.py's names) the AVERAGE over the captured launches of
  dram__bytes_read.sum + dram__bytes_write.sum            -> dram_bytes_per_launch
  l1tex__m_xbar2l1tex_read_bytes.sum + l1tex__m_l1tex2xbar_write_bytes.sum -> l2_bytes_per_launch (L2 <-> SM traffic)
plus selected throughput / stall figures for profiles/*.md.
"""
import csv
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
CLASS = [("cheb_solve_coop_kernel", "spmm_cg"), ("bsr3_cheb_iter_kernel", "spmm_cg"), ("cheb_apply_narrow_kernel", "spmm_cg"),
         ("bsr3_spmm_kernel", "spmm_K"), ("cheb_update_kernel", "cheb_update")]
EXTRA = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct",
         "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
         "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_sample_buffers"]


def main():
    raw, l2file = sys.argv[1], sys.argv[2]
    ref = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    path = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        if name not in ix or r[ix[name]] in ("", "n/a"):
            return None
        return float(r[ix[name]].replace(",", "")) * UNIT.get(units[ix[name]], 1.0)

    acc = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        cls = next((c for k, c in CLASS if k in name), None)
        if cls is None:
            continue
        short = name.split("(")[0].replace("void ", "")
        e = acc.setdefault((cls, short), {"n": 0, "dram": 0.0, "l2": 0.0, "extra": {k: 0.0 for k in EXTRA}})
        e["n"] += 1
        e["dram"] += (val(r, "dram__bytes_read.sum") or 0) + (val(r, "dram__bytes_write.sum") or 0)
        e["l2"] += (val(r, "l1tex__m_xbar2l1tex_read_bytes.sum") or 0) + (val(r, "l1tex__m_l1tex2xbar_write_bytes.sum") or 0)
        for k in EXTRA:
            e["extra"][k] += val(r, k) or 0.0
    l2 = json.load(open(l2file))
    out = {"ref_level": ref, "solve_path": path, "source_capture": os.path.basename(raw),
           "l2_peak_gbs": max(l2["l2_read_gbs_32MB"], l2["l2_read_gbs_64MB"]), "l2_peak_source": "profiles/micro/l2_bandwidth.cu on the same box",
           "hbm_read_gbs_same_kernel": l2["hbm_read_gbs_2GB"], "kernels": {}}
    for (cls, short), e in acc.items():
        ent = {"class": cls, "launches_captured": e["n"], "dram_bytes_per_launch": e["dram"] / e["n"], "l2_bytes_per_launch": e["l2"] / e["n"],
               "source": f"ncu --set full, average over {e['n']} launches ({os.path.basename(raw)})"}
        ent.update({k: v / e["n"] for k, v in e["extra"].items()})
        out["kernels"][short] = ent
        # bench.py looks the dominant class up by name; the per-panel (widest) variant of a class wins
        if cls not in out or ent["dram_bytes_per_launch"] > out[cls]["dram_bytes_per_launch"]:
            out[cls] = {k: ent[k] for k in ("dram_bytes_per_launch", "l2_bytes_per_launch", "source")}
            out[cls]["kernel"] = short
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dominant_kernel_traffic.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}, indent=1))


if __name__ == "__main__":
    main()
