"""ONE ensemble sharded over the GPUs of a box (BASELINE configs[3]: 5 folds x 4 modalities x 24 seeds = 480
cVAEs on 1 / 2 / 4 / 8 B200).

Members are independent, so rank r trains and scores only the members `distributed.shard_members` deals to it
(cost-sorted round-robin: every rank gets the same D=116 : D=348 mix) with NO collective while training.  The single
exchange is one all-gather (NCCL over NVLink; gloo in the CPU tests) of a fixed-size float64 record per member

    { subject AUC | per-ROI mean (D_max) | per-ROI std (D_max) | per-ROI AUC (D_max) | per-subject deviation (N_test_max) }

after which every rank holds the whole table and can average the modalities of a fold that were scored on different
GPUs (multimodal_kfold_cvae_group_analysis_1x1.py:212-215) and compute the fold-level ROC-AUC (:105-157).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from . import distributed as nd
from . import scoring, workloads
from .ensemble import EnsembleTrainer


@dataclass
class GatheredScores:
    table: torch.Tensor          # [n_members, 1 + 3 * d_max + n_test_max] float64, identical on every rank
    d_max: int
    n_test_max: int
    grid: list                   # (fold, modality name, seed) per member, global order

    def subject_auc(self):
        return self.table[:, 0]

    def roi_stats(self, i, d):
        m = self.d_max
        return self.table[i, 1:1 + d], self.table[i, 1 + m:1 + m + d], self.table[i, 1 + 2 * m:1 + 2 * m + d]

    def subject_deviation(self, i, n):
        o = 1 + 3 * self.d_max
        return self.table[i, o:o + n]


def fold_seed_average(gs: GatheredScores, n_test: Sequence[int]):
    """Mean over the modalities of every (fold, seed) of the per-subject deviation (group analysis :212-215).
    Pure tensor code (runs on CPU tensors under gloo too).  Returns {(fold, seed): [N_test_fold] float64}."""
    groups = {}
    for i, (f, _, s) in enumerate(gs.grid):
        groups.setdefault((f, s), []).append(i)
    out = {}
    for key, idx in groups.items():
        n = int(n_test[idx[0]])
        out[key] = torch.stack([gs.subject_deviation(i, n) for i in idx]).mean(0)
    return out


class EnsembleRunner:
    """Train + score one (fold x modality x seed) ensemble, sharded over torch.distributed ranks."""

    def __init__(self, hw: workloads.HostWorkload, n_seeds: int, device, seed0: int = 0, pin: bool = False,
                 score_mode: str = "sample", rank: int = None, world: int = None):
        """rank / world default to the torch.distributed group; world=1 (explicit) = the whole ensemble on this
        GPU with no collective (also inside a distributed run: the unsharded reference of the N=1 vs N>1 test)."""
        self.rank, self.world = nd.world() if world is None else (int(rank or 0), int(world))
        self.hw, self.n_seeds, self.device = hw, n_seeds, torch.device(device)
        self.grid = [(f, name, seed0 + s) for f in range(len(hw.folds)) for name in hw.names for s in range(n_seeds)]
        cost = [workloads.train_flops_per_sample(hw.dims[name], hw.c_dim, hw.hidden, hw.latent) for _, name, _ in self.grid]
        self.owned: List[int] = nd.shard_members(len(self.grid), self.rank, self.world, cost)
        self.wl = workloads.to_device(hw, self.device, n_seeds=n_seeds, seed0=seed0, members=self.owned, pin=pin)
        self.trainer = EnsembleTrainer(self.wl.specs, device=self.device)
        # the reference's test script decodes a SAMPLED z (cVAE_multimodal.pred_recon, cVAE.py:1198-1208) for the test
        # rows; the normative statistics of the training rows use the same mode
        self.scorer = scoring.DeviationScorer(self.trainer, [s.xc for s in self.wl.specs], self.wl.test_xc,
                                              self.wl.train_hc_mask, self.wl.test_labels, mode=score_mode,
                                              params_untouched=True)      # the runner owns the trainer: nobody writes its tensors
        self.n_test_all = [hw.folds[f].test_x[name].shape[0] for f, name, _ in self.grid]
        self.d_max = max(hw.dims.values())
        self.n_test_max = max(self.n_test_all)

    # ---- training: no collective ---------------------------------------------------------------------------
    def train_epochs(self, epochs: int, record_losses: bool = False):
        return self.trainer.train_epochs(epochs, record_losses=record_losses)

    def train_steps(self, n_steps: int, record_losses: bool = False):
        return self.trainer.train_steps(n_steps, record_losses=record_losses)

    # ---- scoring + the one exchange step ------------------------------------------------------------------------
    def local_records(self) -> torch.Tensor:
        """[len(owned), 1 + 3 d_max + n_test_max] float64 records of this rank's members (one segment per member)."""
        return self.scorer.member_table(d_max=self.d_max, n_test_max=self.n_test_max)      # one launch

    def score(self) -> GatheredScores:
        self.scorer.run()
        table = nd.gather_member_tables(self.local_records(), self.owned, len(self.grid), collective=self.world > 1)
        return GatheredScores(table=table, d_max=self.d_max, n_test_max=self.n_test_max, grid=self.grid)

    def fold_auc(self, gs: GatheredScores):
        """ROC-AUC of the modality-averaged deviation per (fold, seed) -- the number the group analysis reports."""
        avg = fold_seed_average(gs, self.n_test_all)
        keys = sorted(avg)
        labels = [torch.from_numpy((self.hw.folds[f].test_df["DIA"].to_numpy() != self.hw.hc_label).astype(np.uint8)).to(self.device)
                  for f, _ in keys]
        aucs = scoring.auc([avg[k].float() for k in keys], labels)
        return {k: float(a[0]) for k, a in zip(keys, aucs)}

    def close(self):
        self.trainer.close()
