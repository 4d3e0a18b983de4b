"""Dev tool: turn the outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.
Run in the build container: needs `ncu` to read the .ncu-rep."""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
G = "gpurun_out"
for page in ("raw", "source"):
    with open("%s/r02_tcp_%s.csv" % (G, "raw" if page == "raw" else "src"), "w") as f:
        subprocess.run(["ncu", "-i", G + "/r02_tcp.ncu-rep", "--page", page, "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
body = subprocess.run([sys.executable, "profiles/ncu_summary.py", G + "/r02_tcp_raw.csv", G + "/r02_tcp_src.csv"],
                      capture_output=True, text=True, check=True).stdout
hdr = """ncu --set full --clock-control none --import-source on -k regex:train_tcp_kernelILb0 -s 3 -c 1 : python tools/run_tcp.py 24 20 5
(nmb::tcp::train_tcp_kernel<false>, 480 members (cfg4) x 20 minibatch steps = one bench.py step, dealt as 2 400 work items of 4 steps;
 round-2 kernel: two-part MMA steps (42 instead of 62 MMA steps per D=116 minibatch step), forward-only instantiation split off)
"""
rows = list(csv.reader(open(G + "/r02_tcp_raw.csv")))
d = dict(zip(rows[0], rows[2]))
extra = "\n-- tensor pipe / memory\n"
for k in ["sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
          "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "launch__block_size",
          "launch__shared_mem_per_block_dynamic"]:
    if k in d:
        extra += "   %s = %s\n" % (k, d[k])
t, rd, wr = float(d["gpu__time_duration.sum"]), float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
rows2 = list(csv.reader(open(G + "/r02_launches.csv")))
hi = [i for i, r in enumerate(rows2) if r and r[0] == "ID"][0]
H, data = rows2[hi], rows2[hi + 1:]
ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
agg = collections.defaultdict(list)
for r in data:
    if len(r) > vi and r[mi] == "gpu__time_duration.sum":
        agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
per = lambda name: [sum(v) / len(v) / 1e6 for k, v in agg.items() if name in k][0]
mv, tk, xp = per("tcp_move"), per("train_tcp"), per("xprep")
share = tk / (tk + 2 * mv + xp)
bench = json.loads(open(G + "/bench_r02.json").read().strip().splitlines()[-1])
reading = """
reading: %.2f ms under ncu for the persistent kernel alone (bench: %.2f ms per training call including xprep %.2f ms and the two
state-conversion launches, %.2f ms each on average; the persistent kernel is %.2f of the call).  DRAM %.1f GB / launch = %.1f MB per
member-step at %s %% of peak DRAM throughput (Adam state r/w 1.44 MB + BF16 weight planes + activation stash; L2 sector hit rate %s %%:
the 148 resident working sets exceed the L2).  Tensor pipe %s %% of active cycles (x3 BF16 passes), its shared-memory operand reads
%s %% of elapsed.  About 70 %% of the executed instructions are the bounded spin-waits of the roles that are waiting
(long_scoreboard = mbarrier try_wait): the launch is bound by the dependent chain epilogue -> MMA issue -> tile -> MMA -> epilogue
of 33 (D=116) / 46 (D=348) items per minibatch step, see r02_tcp_timeline_d116.txt / _d348.txt.
""" % (t, bench["roofline"]["kernel_ms"], xp, mv, share, rd + wr, (rd + wr) * 1e3 / 9600,
       d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][:4], d["lts__t_sector_hit_rate.pct"][:4],
       d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][:4],
       d["sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][:4])
open("profiles/r02_tcp_ncu_summary.txt", "w").write(hdr + body + extra + reading)
tot = sum(sum(v) for v in agg.values())
out = """ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tcp_|xprep|recon|auc_kernel|deviation_kernel|stats_kernel|pack_rows" :
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline
(every libnmb launch of the default bench run incl. the e2e leg and the deviation scoring; cold-cache, serialised: compare SHARES)  unit=ns
share = of all libnmb kernels

"""
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out += "%-62s launches=%4d total_ms=%9.3f per_launch_ms=%8.4f share=%.3f\n" % (k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6, sum(v) / tot)
out += """
training call = xprep + tcp_move (rows -> lane-major state, weight planes) + train_tcp_kernel + tcp_move (back):
train_tcp_kernel is %.2f of it; bench.py's CUDA events bracket all four launches (%.2f ms per call live).
""" % (share, bench["roofline"]["kernel_ms"])
open("profiles/r02_launch_list_summary.txt", "w").write(out)
shutil.copy(G + "/bench_r02.json", "profiles/r02_bench_line.json")
tl = ("tools/trace_tcp.py 1 33 42  (20 members, one per CTA, launch-local minibatch step 2 of CTA 0 = a D=116 member; %globaltimer, "
      "microseconds)\ncolumns: item  arrive  start(after accumulator barrier)  end(published) | wait  work\nE = epilogue group 0 "
      "(rows 0..127), G = group 1 (rows 128..255), A = optimiser group (shared items appear in all three),\nS = MMA steps (deps ready, "
      "tiles landed, issued), P = producer (deps ready, issued)\n\n")
open("profiles/r02_tcp_timeline_d116.txt", "w").write(tl + open(G + "/trace_r02_d116.txt").read())
open("profiles/r02_tcp_timeline_d348.txt", "w").write(tl.replace("1 33 42", "1 46 84 big").replace("D=116", "D=348 (early fusion)") +
                                                       open(G + "/trace_r02_d348.txt").read())
import hashlib
src_hash = hashlib.sha256(b"".join(open(os.path.join("multi_modal_normative_modeling_b200", "csrc", f), "rb").read()
                                   for f in ("nmb_train_tcp.cu", "nmb_tcp.h"))).hexdigest()
json.dump({"kernel": "nmb::tcp::train_tcp_kernel", "kernel_source_sha256": src_hash,
           "note": "bench.py reports this figure only while the kernel sources hash to kernel_source_sha256",
           "capture": "profiles/r02_tcp_ncu_summary.txt (ncu --set full, 480 members x 20 minibatch steps = one bench step)",
           "dram_bytes_read_per_launch": rd * 1e9, "dram_bytes_write_per_launch": wr * 1e9,
           "dram_bytes_per_launch": (rd + wr) * 1e9, "gpu_time_ms_under_ncu": t}, open("profiles/train_kernel_traffic.json", "w"), indent=1)
# the forward-only (reconstruction) instantiation of the same kernel and the scoring breakdown
if os.path.exists(G + "/r02_recon.ncu-rep"):
    for page in ("raw", "source"):
        with open("%s/r02_recon_%s.csv" % (G, "raw" if page == "raw" else "src"), "w") as f:
            subprocess.run(["ncu", "-i", G + "/r02_recon.ncu-rep", "--page", page, "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
    rbody = subprocess.run([sys.executable, "profiles/ncu_summary.py", G + "/r02_recon_raw.csv", G + "/r02_recon_src.csv"],
                           capture_output=True, text=True, check=True).stdout
    rr = list(csv.reader(open(G + "/r02_recon_raw.csv")))
    rd_ = dict(zip(rr[0], rr[2]))
    rextra = "\n-- tensor pipe / memory\n"
    for k in ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size"]:
        if k in rd_:
            rextra += "   %s = %s\n" % (k, rd_[k])
    open("profiles/r02_recon_ncu_summary.txt", "w").write(
        "ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:train_tcp_kernelILb1 -s 4 -c 1 : "
        "python tools/time_scoring.py\n(nmb::tcp::train_tcp_kernel<true>: the forward-only program of nmb_ensemble_reconstruct, one launch "
        "over the training rows of the 480 cfg4 members = 1 920 tiles of 256 rows dealt as 960 work items)\n" + rbody + rextra)
if os.path.exists(G + "/scoring_r02.txt"):
    open("profiles/r02_scoring_breakdown.txt", "w").write(
        "tools/time_scoring.py (cfg4, 480 members: 384 000 training rows + 96 000 test rows per pass; CUDA events, 10 repetitions)\n" +
        "".join(l for l in open(G + "/scoring_r02.txt") if " gpu " in l))
print(reading)
print("share %.3f  value %.1f M samples/s  e2e %.1f M  deviation %.1f M subjects/s" %
      (share, bench["value"] / 1e6, bench["e2e"]["value"] / 1e6, bench["deviation"]["value"] / 1e6))
