"""Warp-instructions per unit of the hot kernels: runs tools/instr_probe.py under ncu (instruction and pipe counters
only), matches the launches to the probe's unit counts and writes profiles/instr_counts_r2.json (read by bench.py for
the issue rooflines).  Run on the GPU box:  python tools/instr_counts.py"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = os.path.join(ROOT, "gpurun_out", "instr_probe_ncu.csv")
metrics = ("smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,"
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,"
           "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum")
if "--parse-only" not in sys.argv:      # re-read an existing gpurun_out/instr_probe_ncu.csv (no GPU needed)
  subprocess.check_call(["ncu", "--metrics", metrics, "--clock-control", "none", "--csv", "--log-file", log,
                       "-k", "regex:k_polymul_w32|k_ntt_w32|k_exact_w32|k_matvec|k_cdf_aes|k_cdf_chacha|k_stream_seq|k_ber_lanes",
                       sys.executable, os.path.join(ROOT, "tools", "instr_probe.py")])
units = json.load(open(os.path.join(ROOT, "gpurun_out", "instr_probe_units.json")))
rows = [r for r in csv.reader(open(log)) if len(r) > 10]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
launches = {}           # ID -> {name, metrics}
for r in rows[1:]:
    lid = r[col["ID"]]
    d = launches.setdefault(lid, {"name": r[col["Kernel Name"]], "m": {}})
    d["m"][r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
ordered = [launches[k] for k in sorted(launches, key=int)]
out = {}
it = iter(ordered)
for kern, key, n in units:
    for L in it:
        if kern in L["name"]:
            m = L["m"]
            out[key] = {"kernel": L["name"][:120], "units": n, "warp_instr": m["smsp__inst_executed.sum"],
                        "per_unit": m["smsp__inst_executed.sum"] / n,
                        "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "alu_pipe_pct": m.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                        "fma_pipe_pct": m.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                        "fmaheavy_pipe_pct": m.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                        "dram_bytes_per_unit": (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)) / n,
                        "ns_under_ncu": m.get("gpu__time_duration.sum"),
                        "source": "profiles/instr_counts_r2.json: ncu counters of tools/instr_probe.py (tools/instr_counts.py)"}
            break
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "instr_counts_r2.json"), "w"), indent=1)
for k, v in out.items():
    print("%-34s %10.1f warp-instr / unit   issue %5.1f %%   alu %5.1f %%   %6.0f B dram / unit" % (k, v["per_unit"], v["issue_active_pct"] or 0, v["alu_pipe_pct"] or 0, v["dram_bytes_per_unit"]))
