"""Summarise an `ncu --page raw --csv` + `--page source --csv` export pair (run in the build container)."""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
for k in keys:
    if k in d:
        print(f"{k} = {d[k]} {u[k]}")
print("-- warp stall reasons (avg warps stalled per issue-active cycle)")
st = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[k]) for k in hdr
      if "average_warps_issue_stalled" in k and k.endswith(".ratio")}
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
    print(f"   {k:22s} {v:.3f}")
rows = list(csv.reader(open(src)))
h, data = rows[1], rows[2:]
si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
tot = sum(int(r[si]) for r in data); te = sum(int(r[ie]) for r in data)
agg, agge = collections.Counter(), collections.Counter()
for r in data:
    t = r[so].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    agg[op] += int(r[si]); agge[op] += int(r[ie])
print(f"-- SASS: {len(data)} instructions; opcode share of (stall samples, executed instructions)")
for op, c in agg.most_common(12):
    print(f"   {op:8s} {c / tot:.3f} {agge[op] / te:.3f}")
