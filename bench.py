#!/usr/bin/env python
"""Benchmark of the cVAE-ensemble hot path (BASELINE.json metric: cVAE train samples/s (all
models) & deviation subjects/s at 1/2/4/8 B200 vs CPU).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, libnmb.so)
    python bench.py --impl reference --steps K --warmup W    # the reference's PyTorch CPU path

Workload (config.workload = "cfg4"): 5 folds x {T1w, T2w, fMRI (D=116), early fusion (D=348)} x 24
seeds = 480 independent cVAEs per GPU (weak scaling: every rank trains its own 480-member
ensemble with distinct seeds), hidden [110,110], latent 10, C=29, batch 256, 800 bootstrap rows
per fold, synthetic HCP-shaped data, reference-exact random initialisation.
One "step" = `--epochs-per-step` epochs (default 5 = 20 minibatch steps) of EVERY member in ONE
fused kernel launch (forward + loss + backward + Adam).  `value` counts training samples only;
deviation scoring is timed separately and reported under "deviation".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cVAE train samples/s (all models)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seeds", type=int, default=24, help="seeds per (fold, modality): 24 -> 480 members per GPU")
    ap.add_argument("--epochs-per-step", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-deviation", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the bounded CPU sample")
    return ap.parse_args()


def config(args, extra=None):
    c = {"workload": "cfg4: 5 folds x 4 modalities (D=116,116,116,348) x %d seeds = %d cVAEs per GPU"
                     % (args.seeds, 20 * args.seeds),
         "hidden": [110, 110], "latent": 10, "c_dim": 29, "batch": 256, "n_train": 800, "n_test": 200,
         "epochs_per_step": args.epochs_per_step, "members_per_gpu": 20 * args.seeds,
         "parallelism": "member-sharded, no gradient collective",
         "l2": "working set (params+Adam+scratch > 600 MB) exceeds the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's eager PyTorch loop, restated in oracle/cvae_torch.py (the reference
# is Python and cannot travel to the GPU box; kind = "port").
def cpu_reference_sample(hw, epochs, budget_s, threads=None):
    """Reference training loop on the host for fold 0 x all 4 modalities (keeps the 3:1 D=116 : D=348 mix).
    With a finite `budget_s` the number of epochs is calibrated so that the sample takes about that long."""
    import torch
    from oracle import cvae_torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    fold = hw.folds[0]

    def one_pass(n_epochs):
        samples, t_total = 0, 0.0
        for name in hw.names:
            d = hw.dims[name]
            torch.manual_seed(42)
            model = cvae_torch.OracleCVAEMultimodal([d], list(hw.hidden), hw.latent, hw.c_dim, 1e-4, 1, True)
            x = torch.from_numpy(fold.train_x[name])
            c = torch.from_numpy(fold.train_c).long()          # int64 one-hots (utils_vae.py:24)
            t0 = time.perf_counter()
            cvae_torch.reference_train_loop(model, [x], [c], "gPoE", n_epochs, hw.batch)
            t_total += time.perf_counter() - t0
            samples += x.shape[0] * n_epochs
        return samples, t_total

    if budget_s < 1e8:
        s0, t0 = one_pass(2)                                   # calibration (also warms the thread pool)
        epochs = max(epochs, int(budget_s / max(t0 / 2, 1e-6)))
    samples, t_total = one_pass(epochs)
    return samples / t_total, threads, len(hw.names), t_total, epochs


def cpu_deviation_sample(hw, threads):
    """pred_recon + deviation + roc/auc on the host for one fold x 4 modalities (subjects/s)."""
    import numpy as np
    import torch
    from oracle import cvae_torch, deviation as odev
    torch.set_num_threads(threads)
    fold = hw.folds[0]
    labels = (fold.test_df["DIA"].to_numpy() != hw.hc_label).astype(np.int64)
    n, t_total = 0, 0.0
    for name in hw.names:
        d = hw.dims[name]
        torch.manual_seed(42)
        model = cvae_torch.OracleCVAEMultimodal([d], list(hw.hidden), hw.latent, hw.c_dim, 1e-4, 1, True)
        x = torch.from_numpy(fold.test_x[name]); c = torch.from_numpy(fold.test_c).long()
        xt = torch.from_numpy(fold.train_x[name]); ct = torch.from_numpy(fold.train_c).long()
        t0 = time.perf_counter()
        pred = model.pred_recon([x], c, "gPoE")[0].numpy()
        pred_tr = model.pred_recon([xt], ct, "gPoE")[0].numpy()
        roi = odev.recon_deviation_roi(fold.test_x64[name], pred)
        subj = odev.recon_deviation(fold.test_x64[name], pred)
        mean, std = odev.normative_stats(odev.recon_deviation_roi(fold.train_x[name], pred_tr))
        z = odev.zscores(roi, mean, std)
        _ = [odev.auc(z[:, j], labels) for j in range(d)]
        _ = odev.auc(subj, labels)
        t_total += time.perf_counter() - t0
        n += x.shape[0]
    return n / t_total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from multi_modal_normative_modeling_b200 import workloads
    hw = workloads.build_host_workload()
    threads = os.cpu_count() or 1
    # each step = a bounded sample (1 fold x 4 modalities x epochs_per_step epochs = 16k samples)
    for _ in range(args.warmup):
        cpu_reference_sample(hw, 1, 1e9, threads)
    t0 = time.perf_counter()
    samples = 0
    for _ in range(args.steps):
        rate, _, models, dt, _ = cpu_reference_sample(hw, args.epochs_per_step, 1e9, threads)
        samples += rate * dt
    el = time.perf_counter() - t0
    value = samples / el
    sample = "1 fold x 4 modalities x %d epochs per step, oracle/cvae_torch.py eager loop" % args.epochs_per_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.armed = False          # NVML is initialised during the warm-up; samples count only inside the timed region
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            self.ready.set()
            while not self.stop_flag:
                if self.armed:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                time.sleep(0.005)
        except Exception as e:          # NVML unavailable: report that instead of inventing numbers
            self.reasons.add("nvml_error:%s" % type(e).__name__)
            self.ready.set()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from multi_modal_normative_modeling_b200 import EnsembleTrainer, pack_rows, scoring, workloads
    from multi_modal_normative_modeling_b200 import distributed as nd

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    hw = workloads.build_host_workload()
    wl = workloads.to_device(hw, dev, n_seeds=args.seeds, seed0=rank * args.seeds, pin=True)
    tr = EnsembleTrainer(wl.specs, device=dev)
    spe = tr.steps_per_epoch[0]
    n_steps = args.epochs_per_step * spe
    samples_per_step = wl.samples_per_epoch * args.epochs_per_step
    flops_per_step = wl.flops_per_epoch * args.epochs_per_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) ------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        tr.train_steps(n_steps)
    sampler.ready.wait(5.0)
    barrier()
    sampler.armed = True
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = tr.gpu_launches
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for a, b in ev:
        a.record()
        tr.train_steps(n_steps)
        b.record()
    t_end.record()
    barrier()
    sampler.armed = False                       # re-armed for the end-to-end timed region below
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    launches = tr.gpu_launches - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t[0])
    value = world * samples_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers ----------------------------------
    keys = list(wl.host_buffers)
    h2d = sum(x.numel() * 4 + c.numel() * 4 for x, c in wl.host_buffers.values())
    loss_host = torch.empty((tr.n, n_steps, 3), dtype=torch.float32).pin_memory()

    # Every step: pinned host -> device copy of every dataset, re-pack, train, losses back to pinned host memory.
    # The copies of step i + 1 travel on a side stream (two staging sets) while step i trains; the losses of step i
    # leave on a third stream.  All of it is inside the timed region.
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage = [{k: (torch.empty_like(x, device=dev), torch.empty_like(c, device=dev)) for k, (x, c) in wl.host_buffers.items()}
             for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]        # staging set filled
    ev_free = [torch.cuda.Event() for _ in range(2)]      # staging set consumed by pack_rows
    ev_loss = torch.cuda.Event()

    def upload(slot):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[slot])
            for k in keys:
                x, c = wl.host_buffers[k]
                stage[slot][k][0].copy_(x, non_blocking=True)
                stage[slot][k][1].copy_(c, non_blocking=True)
            ev_in[slot].record(s_in)

    def e2e_run(n):
        for e in ev_free:
            e.record(main)
        upload(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                upload(slot ^ 1)
            main.wait_event(ev_in[slot])
            for k in keys:
                pack_rows(stage[slot][k][0], stage[slot][k][1], out=wl.packed[k])
            ev_free[slot].record(main)
            s_out.wait_event(ev_loss)                     # previous losses have left before the buffer is rewritten
            losses = tr.train_steps(n_steps, record_losses=True)
            ev_loss.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_loss)
                loss_host.copy_(losses, non_blocking=True)
                losses.record_stream(s_out)
        main.wait_stream(s_out)
    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.armed = True
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = world * samples_per_step * args.steps / (e2e_ms * 1e-3)
    final_loss = float(loss_host[:, -1, 0].mean())

    # ---- deviation scoring: reconstruct -> normative stats -> deviation/z -> AUC -> all-gather --
    deviation = None
    if not args.no_deviation:
        scorer = scoring.DeviationScorer(tr, [s.xc for s in wl.specs], wl.test_xc, wl.train_hc_mask, wl.test_labels)

        def dev_step():
            # 6 libnmb launches (2 x reconstruct, stats, deviation / z, 2 x AUC), then one fixed-size record per
            # member {subject AUC | per-ROI mean | std | AUC} and the all-gather (NCCL over NVLink when N > 1)
            scorer.run()
            rec = scorer.member_records()
            owned = list(range(rank * tr.n, (rank + 1) * tr.n))
            table = nd.gather_member_tables(rec, owned, world * tr.n)
            return table, scorer.auc_subj
        for _ in range(max(args.warmup, 5)):
            dev_step()
        barrier()
        reps = max(5, min(args.steps, 10))
        dev_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in dev_ev:
            a.record()
            table, subj_auc = dev_step()
            b.record()
        barrier()
        # per-pass device time, median over the passes (max over ranks): a pass is a handful of short launches, so a
        # single host hiccup (first process on a fresh box: lazy module loading, clock ramp) would otherwise dominate
        pass_ms = sorted(a.elapsed_time(b) for a, b in dev_ev)
        dms = pass_ms[len(pass_ms) // 2] * reps
        dev_worst = pass_ms[-1]
        if world > 1:
            t = torch.tensor([dms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dms = float(t[0])
        # algorithmic bytes of the streaming deviation kernel: read x (ldx) + xhat, write roi + z + subj
        dev_bytes = sum(t[0].shape[0] * (4 * (3 * s.input_dims[0] + t[0].shape[1]) + 4)
                        for t, s in zip(wl.test_xc, wl.specs))
        deviation = {"value": world * wl.test_subjects * reps / (dms * 1e-3), "unit": "subjects/s",
                     "ms_per_pass": dms / reps, "ms_worst_pass": dev_worst, "timing": "median pass of %d, CUDA events" % reps,
                     "subjects_per_pass": world * wl.test_subjects,
                     "mean_subject_auc": float(subj_auc.mean()),
                     "gathered_table": list(table.shape), "launches_per_pass": 6,
                     "streaming_kernel_algorithmic_bytes": dev_bytes}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
        achieved = flops_per_step / (kernel_ms * 1e-3) / 1e12
        clocks = sampler.summary()
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "train_kernel_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        engine = tr.engine()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3->f32" if engine != "fp32" else "f32",
                "data": "synthetic",
                "config": config(args, {"final_mean_total_loss": final_loss, "engine": engine}),
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": achieved / peak_tf, "traffic": traffic,
                             "kernel": "nmb::tcp::train_tcp_kernel" if engine == "tcgen05-pipelined" else "nmb::train_kernel",
                             "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": flops_per_step,
                             "peak_source": peak_src,
                             "kernel_ms_note": "CUDA events bracket the four launches of a training call (xprep, state "
                                               "conversion in, persistent kernel, state conversion out); the persistent kernel is "
                                               "0.92 of it (profiles/r01b_launch_list_summary.txt)",
                             "hbm": None if not traffic else {
                                 "achieved_GBps": traffic / (kernel_ms * 1e-3) / 1e9, "peak_GBps": peaks.get("hbm_gbs"),
                                 "frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                 "note": "measured DRAM bytes per launch (ncu, profiles/train_kernel_traffic.json) over the "
                                         "live launch time: the kernel is latency-bound between its two roofs"},
                             "note": "achieved = algorithmic (FP32-equivalent) FLOPs; every product is executed as 3 BF16 "
                                     "tcgen05 passes (hi*hi, lo*hi, hi*lo, FP32 accumulate) to meet the 1e-4 parity bar, so "
                                     "the tensor pipe executes 3x these FLOPs: executed_frac = %.4f of the measured peak"
                                     % (3 * achieved / peak_tf),
                             "executed_tensor_frac": 3 * achieved / peak_tf},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": loss_host.numel() * 4, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "clocks": clocks}
        if deviation:
            line["deviation"] = deviation
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, cores, models, dt, ep = cpu_reference_sample(hw, 5, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "fold 0 x %d modalities (3 x D=116 + D=348) x %d epochs, %.1f s, "
                                              "oracle/cvae_torch.py eager PyTorch loop" % (models, ep, dt),
                                    "deviation_subjects_per_s": cpu_deviation_sample(hw, threads)}
        print(json.dumps(line))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
