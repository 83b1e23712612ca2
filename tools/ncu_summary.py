#!/usr/bin/env python
"""Summarise an `ncu --set full` report of tools/profile_target.py: one row per backbone kernel of the LAST forward
(time, registers, pipes, DRAM bytes, shared-memory wavefronts) + a JSON of DRAM bytes per launch that bench.py reads as
roofline.traffic.  Usage: ncu_summary.py report.ncu-rep summary.txt traffic.json [batch] [size] [all]
("all": one row per kernel launch in the report, whatever it is -- used for the head / post-processing kernels of a bench step)"""
import csv
import json
import subprocess
import sys

rep, out_txt, out_json = sys.argv[1:4]
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
size = int(sys.argv[5]) if len(sys.argv) > 5 else 96

WANT = [("time us", ("gpu__time_duration.sum",)),
        ("regs", ("launch__registers_per_thread",)),
        ("threads", ("launch__block_size",)),
        ("warps active %", ("sm__warps_active.avg.pct_of_peak_sustained_active",)),
        ("issue active %", ("sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
                            "smsp__issue_active.avg.pct_of_peak_sustained_active")),
        ("fma pipe %", ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")),
        ("tensor pipe %", ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                           "sm__pipe_tensor_subpipe_tmem_cycles_active.avg.pct_of_peak_sustained_active",
                           "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")),
        ("dram %", ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed")),
        ("dram read", ("dram__bytes_read.sum",)),
        ("dram write", ("dram__bytes_write.sum",)),
        ("smem wavefronts %", ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
                               "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")),
        ("warp instr", ("smsp__inst_executed.sum", "sm__inst_executed.sum"))]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def load(path):
    """one dict per kernel launch of a report: {"name": kernel name, label: (value, unit) for the labels of WANT the report holds}"""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in body:
        if "Kernel Name" not in col or len(r) <= col["Kernel Name"]:
            continue
        d = {"name": r[col["Kernel Name"]]}
        for label, names in WANT:
            for n in names:
                if n in col and r[col[n]] != "":
                    d[label] = (r[col[n]], units[col[n]])
                    break
        out.append(d)
    return out


# several reports (comma separated) are concatenated in order: e.g. the stem captured on its own (its 928 threads x 64 registers
# leave no room for the instrumented passes of --set full) followed by the block kernels
launches = [d for path in rep.split(",") for d in load(path)]
if len(sys.argv) > 6 and sys.argv[6] == "all":
    last = launches
    names = [f"#{i}" for i in range(len(last))]
else:
    backbone = [d for d in launches if "stem" in d["name"] or "blaze_block" in d["name"] or "blaze_chain" in d["name"]]
    # the last forward: stem, blocks 0-5, then either one kernel per block or the two chain kernels (6-10 + tail 11, 12-15);
    # names follow bench.py's merged rows
    n_chain = sum(1 for r in backbone if "blaze_chain" in r["name"])
    if n_chain:
        last = backbone[-9:]
        tail = any("blaze_block" in r["name"] for r in last[8:9]) is False and len([r for r in last if "blaze_chain" in r["name"]]) == 2
        names = ["stem"] + [f"block{i}" for i in range(6)] + ["blocks6-11", "blocks12-15"]
        if sum(1 for r in backbone[-10:] if "blaze_chain" in r["name"]) == 2 and "blaze_chain" not in backbone[-2]["name"]:
            last = backbone[-10:]             # chains without the tail: block 11 is its own kernel between them
            names = ["stem"] + [f"block{i}" for i in range(6)] + ["blocks6-10", "block11", "blocks12-15"]
    else:
        last = backbone[-17:]
        names = ["stem"] + [f"block{i}" for i in range(16)]
traffic = {}
with open(out_txt, "w") as f:
    f.write(f"# ncu --clock-control none (--set full; the stem without the instrumented sections), tools/profile_target.py {size} {batch}; one row per backbone kernel\n")
    f.write("kernel | " + " | ".join(n for n, _ in WANT) + "\n")
    for nm, r in zip(names, last):
        vals = []
        for n, _ in WANT:
            if n not in r:
                vals.append("n/a")
            elif n.startswith("dram r") or n.startswith("dram w"):
                vals.append(f"{to_bytes(*r[n]) / 1e9:.6f} Gbyte")
            else:
                vals.append(r[n][0])
        kn = r["name"].replace("void ", "").replace("<unnamed>::", "")
        f.write(f"{nm} {kn[:60]} | " + " | ".join(vals) + "\n")
        if "dram read" in r and "dram write" in r:
            traffic[nm] = {"dram_read_bytes": to_bytes(*r["dram read"]), "dram_write_bytes": to_bytes(*r["dram write"]), "batch": batch, "size": size}
json.dump(traffic, open(out_json, "w"), indent=1)
print(open(out_txt).read())
