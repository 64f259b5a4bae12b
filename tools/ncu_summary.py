#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per captured launch the kernel name, duration, DRAM bytes, L2->SM bytes,
tensor-pipe activity, issue utilisation, registers.  `--json out.json` also writes the mean DRAM traffic per launch
that bench.py reports as roofline.traffic.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json profiles/r02_conv_traffic.json] [--md profiles/x.md]
"""
import csv
import io
import json
import subprocess
import sys

UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
WANT = {"gpu__time_duration.sum": "time_s", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct", "launch__registers_per_thread": "regs",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "launch__grid_size": "grid",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "smsp__cycles_elapsed.avg.per_second": "sm_hz"}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        d = {"kernel": r[col["Kernel Name"]].split("(")[0]}
        for m, k in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", ""))
                d[k] = v * UNITS.get(units[col[m]], 1.0)
        if "dram_read" in d:
            d["dram_bytes"] = d["dram_read"] + d.get("dram_write", 0.0)
        out.append(d)
    for d in out:
        print(json.dumps(d))
    if "--json" in sys.argv:
        tr = [d["dram_bytes"] for d in out if "dram_bytes" in d]
        json.dump({"source": rep + " (ncu --set full)", "launches": len(tr),
                   "dram_bytes_per_launch_mean": sum(tr) / max(1, len(tr)), "per_launch": tr,
                   "kernels": sorted({d["kernel"] for d in out})}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
    if "--md" in sys.argv:
        with open(sys.argv[sys.argv.index("--md") + 1], "w") as f:
            f.write(f"# ncu --set full summary of `{rep}`\n\n| kernel | ms | DRAM GB | L2->SM GB | tensor pipe % | issue % | regs | grid |\n|---|---|---|---|---|---|---|---|\n")
            for d in out:
                f.write(f"| {d['kernel'][:60]} | {d.get('time_s', 0) * 1e3:.3f} | {d.get('dram_bytes', 0) / 1e9:.3f} | "
                        f"{d.get('l2_to_sm', 0) / 1e9:.2f} | {d.get('tensor_pct', 0):.1f} | {d.get('issue_pct', 0):.1f} | "
                        f"{int(d.get('regs', 0))} | {int(d.get('grid', 0))} |\n")


if __name__ == "__main__":
    main()
