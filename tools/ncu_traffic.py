"""Reads an `ncu --set full` report (here, without a GPU: `ncu -i <rep> --page raw --csv`) and writes, per kernel of the
library, duration, DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), the throughput figures and occupancy /
registers -- as a markdown table (stdout) and, with --json <path> --points <S>, the per-point DRAM traffic that bench.py
reports as `roofline.traffic` (profiles/r2_ncu_traffic.json).
Usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep [--json profiles/r2_ncu_traffic.json --points 916000]"""
import csv
import io
import json
import re
import subprocess
import sys

NAMES = [("mlp_bwd_pipe_kernel<(bool)1>", "mlp_bwd_hash_scatter"), ("mlp_bwd_pipe_kernel<1>", "mlp_bwd_hash_scatter"),
         ("mlp_bwd_pipe_kernel<(bool)0>", "mlp_bwd"), ("mlp_bwd_pipe_kernel<0>", "mlp_bwd"),
         ("mlp_bwd_pipe_kernel", "mlp_bwd"), ("hash_bwd_kernel", "hash_encode_bwd"), ("hash_fwd_kernel", "hash_encode_fwd"),
         ("mlp_kernel<(bool)0>", "mlp_fwd"), ("mlp_kernel<0>", "mlp_fwd"), ("mlp_kernel<(bool)1>", "mlp_bwd_serial"),
         ("mlp_kernel<1>", "mlp_bwd_serial"), ("mlp_kernel", "mlp_fwd"),
         ("adam_kernel", "adam"), ("composite_fwd_kernel", "composite_fwd"), ("composite_bwd_kernel", "composite_bwd"),
         ("march_expand", "march_expand"), ("march_warp_kernel", "march_count"), ("march_thread_kernel", "march_count")]
COLS = {"gpu__time_duration.sum": "time", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct", "launch__registers_per_thread": "regs",
        "sm__inst_executed.sum": "inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct"}
UNIT_SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):           # already exported on the GPU box (`ncu -i <rep> --page raw --csv`): large reports stay there
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kname = idx["Kernel Name"]
    out = []
    for r in data:
        rec = {"kernel": r[kname]}
        for col, key in COLS.items():
            if col in idx and r[idx[col]] not in ("", "n/a"):
                v = float(r[idx[col]].replace(",", ""))
                rec[key] = v * UNIT_SCALE.get(units[idx[col]], 1.0)
        out.append(rec)
    print("| kernel | ms | DRAM MB (rd+wr) | DRAM % | L1/TEX % | L2 % | SM % | tensor % | issue % | warps % | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for r in out:
        short = re.sub(r"\(.*", "", r["kernel"])[:60]
        g = lambda k, f="{:.1f}": f.format(r[k]) if k in r else "-"
        print(f"| `{short}` | {g('time', '{:.4f}')} | {(r.get('rd', 0) + r.get('wr', 0)) / 1e6:.1f} | {g('dram_pct')} | {g('l1_pct')} | "
              f"{g('l2_pct')} | {g('sm_pct')} | {g('tensor_pct')} | {g('issue_pct')} | {g('occ_pct')} | {g('regs', '{:.0f}')} |")
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        pts = float(sys.argv[sys.argv.index("--points") + 1])
        res = {}
        for r in out:
            if r.get("time", 0) < 0.018:        # the occupancy update's small encoder / MLP launches are not the step's kernels
                continue
            for pat, name in NAMES:
                if pat in r["kernel"]:
                    e = res.setdefault(name, {"n": 0, "bytes": 0.0, "ms": 0.0})
                    e["n"] += 1; e["bytes"] += r.get("rd", 0) + r.get("wr", 0); e["ms"] += r.get("time", 0)
                    break
        js = {k: {"dram_bytes_per_point": v["bytes"] / v["n"] / pts, "launches": v["n"], "ms_per_launch_under_ncu": v["ms"] / v["n"],
                  "points_per_launch": pts, "source": rep.split("/")[-1]} for k, v in res.items()}
        with open(path, "w") as f:
            json.dump(js, f, indent=1, sort_keys=True)
        print(f"wrote {path}")


if __name__ == "__main__":
    main()
