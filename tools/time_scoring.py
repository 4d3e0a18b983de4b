"""Dev tool: where the deviation-scoring pass spends its time (per libnmb call, GPU events vs wall clock)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, scoring, _lib
from multi_modal_normative_modeling_b200 import distributed as nd

dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=24)
tr = EnsembleTrainer(wl.specs, device=dev)
tr.train_steps(8)
sc = scoring.DeviationScorer(tr, [s.xc for s in wl.specs], wl.test_xc, wl.train_hc_mask, wl.test_labels, mode="sample", params_untouched=True)
for _ in range(3):
    sc.run(); sc.member_records()
torch.cuda.synchronize()
def timed(fn, reps=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a.record()
    for _ in range(reps): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return a.elapsed_time(b) / reps, (t1 - t0) * 1e3 / reps, (t2 - t0) * 1e3 / reps
print("run()            gpu %.3f ms  host-issue %.3f ms  wall %.3f ms" % timed(sc.run))
print("member_records() gpu %.3f ms  host-issue %.3f ms  wall %.3f ms" % timed(sc.member_records))
rec = sc.member_records()
owned = list(range(tr.n))
print("gather           gpu %.3f ms  host-issue %.3f ms  wall %.3f ms" % timed(lambda: nd.gather_member_tables(rec, owned, tr.n)))
lib, st = sc.lib, torch.cuda.current_stream().cuda_stream
calls = {
 "reconstruct both": lambda: lib.nmb_ensemble_reconstruct_sets(tr.handle, 2, sc.t_xc_both, sc.t_rows_both, sc.mode, None, sc.t_hat_both, None, None, st),
 "stats": lambda: lib.nmb_normative_stats(sc.n_seg, sc.s_x_tr, sc.s_ldx, sc.s_hat_tr, sc.s_mask, sc.s_ntr, sc.s_d, sc.s_stats, st),
 "deviation": lambda: lib.nmb_deviation(sc.n_seg, sc.s_x_te, sc.s_ldx, sc.s_hat_te, sc.s_stats, sc.s_nte, sc.s_d, sc.s_roi, sc.s_z, sc.s_subj, st),
 "auc roi": lambda: lib.nmb_auc(sc.n_seg, sc.s_z, sc.s_lab, sc.s_nte, sc.s_d, sc.s_auc_roi, None, st),
 "auc subj": lambda: lib.nmb_auc(sc.n_seg, sc.s_subj, sc.s_lab, sc.s_nte, sc.s_one, sc.s_auc_subj, None, st),
}
for k, f in calls.items():
    print("%-18s gpu %.3f ms  host-issue %.3f ms  wall %.3f ms" % ((k,) + timed(f)))
