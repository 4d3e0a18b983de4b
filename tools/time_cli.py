"""Dev tool: wall-clock of a whole CSV -> AUC-file run of the multimodal_kfold_* programs on synthetic HCPimage data
(1 000 subjects, 3 modalities x 116 ROIs, 5 folds, E epochs): train + test + group analysis, with the per-phase
breakdown of the train program -- how much is the fused training launch, how much the host prologue (CSV parsing,
merges), how much the GPU / host scaler + covariate + packing step."""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_modal_normative_modeling_b200 import cli, synthetic

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}
for host, fast, pandas_csv in ((False, False, False), (True, False, False), (False, False, True), (False, True, False)):
    with tempfile.TemporaryDirectory() as root:
        synthetic.write_dataset(root, "HCPimage", n=1000, seed=42)
        ns = dict(dataset_resourse="HCPimage", hz_para_list=[110, 110, 10], combine=None, procedure="SE-gPoE", n_splits=5,
                  epochs=epochs, oversample_percentage=1, model="cVAE_multimodal", single_modality=None,
                  base_learning_rate=1e-4, max_learning_rate=5e-3, training_class="nm", ensemble_seeds=1, nmmlp=False,
                  host_prologue=host, fast_csv=fast, pandas_csv=pandas_csv)
        cli.TIMINGS.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); cli.train_main(argparse.Namespace(**ns), root=root); torch.cuda.synchronize(); t1 = time.perf_counter()
        cli.test_main(argparse.Namespace(**ns), root=root); torch.cuda.synchronize(); t2 = time.perf_counter()
        summary = cli.analysis_main(argparse.Namespace(**ns), root=root); t3 = time.perf_counter()
        out[("host_prologue" if host else "gpu_prologue") + ("+arrow_csv" if fast else "+pandas_csv" if pandas_csv else "+native_csv")] = {
            "train_program_s": t1 - t0, "test_program_s": t2 - t1, "group_analysis_s": t3 - t2, "total_s": t3 - t0,
            "train_phases_s": dict(cli.TIMINGS), "mean_auc": float(summary[0][2][0])}
print(json.dumps({"epochs": epochs, "members": 5, "modalities": 3, **out}, indent=1))
