"""Key numbers of an ncu report (raw page) + instruction mix (source page).
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep [out.json]"""
import csv, json, subprocess, sys, os, tempfile
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio"]
out = {}
for i, h in enumerate(hdr):
    if h in keep or ("issue_stalled" in h and "per_issue_active" in h):
        out[h.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")] = vals[i] + " " + units[i]
for k, v in out.items():
    print("%-70s %s" % (k, v))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
    f.write(src)
mix = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_mix.py"), f.name], capture_output=True, text=True).stdout
print(mix)
if len(sys.argv) > 2:
    open(sys.argv[2].replace(".json", "_mix.txt"), "w").write(mix)
