"""Dev tool: deviation-scoring pass timed after a long training run (the bench.py situation)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_normative_modeling_b200 import EnsembleTrainer, workloads, scoring

dev = torch.device("cuda", 0)
hw = workloads.build_host_workload()
wl = workloads.to_device(hw, dev, n_seeds=24)
tr = EnsembleTrainer(wl.specs, device=dev)
sc = scoring.DeviationScorer(tr, [s.xc for s in wl.specs], wl.test_xc, wl.train_hc_mask, wl.test_labels)
def timed(fn, reps=5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
lib, st = sc.lib, torch.cuda.current_stream().cuda_stream
calls = {
 "reconstruct train": lambda: lib.nmb_ensemble_reconstruct(tr.handle, sc.t_xc_tr, sc.t_rows_tr, sc.mode, None, sc.t_hat_tr, None, None, st),
 "reconstruct test": lambda: lib.nmb_ensemble_reconstruct(tr.handle, sc.t_xc_te, sc.t_rows_te, sc.mode, None, sc.t_hat_te, None, None, st),
 "stats": lambda: lib.nmb_normative_stats(sc.n_seg, sc.s_x_tr, sc.s_ldx, sc.s_hat_tr, sc.s_mask, sc.s_ntr, sc.s_d, sc.s_stats, st),
 "deviation": lambda: lib.nmb_deviation(sc.n_seg, sc.s_x_te, sc.s_ldx, sc.s_hat_te, sc.s_stats, sc.s_nte, sc.s_d, sc.s_roi, sc.s_z, sc.s_subj, st),
 "auc roi": lambda: lib.nmb_auc(sc.n_seg, sc.s_z, sc.s_lab, sc.s_nte, sc.s_d, sc.s_auc_roi, None, st),
 "auc subj": lambda: lib.nmb_auc(sc.n_seg, sc.s_subj, sc.s_lab, sc.s_nte, sc.s_one, sc.s_auc_subj, None, st),
}
for total in (8, 200, 1400):
    while int(tr.steps_done()[0]) < total:
        tr.train_steps(20)
    print("after >= %d steps: run() %.3f ms" % (total, timed(sc.run)))
    for k, f in calls.items():
        print("   %-18s %.3f ms" % (k, timed(f)))
