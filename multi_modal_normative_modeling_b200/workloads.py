"""The BASELINE.json workloads on synthetic HCP-shaped data, built once on
the host with the reference's own fold / scaler / covariate pipeline (pipeline.py) and handed to
the fused ensemble kernels.  Used by bench.py, the CLIs' ensemble mode and the end-to-end tests.

cfg4 (the configuration the headline metric is quoted on): 5 folds x {T1w_sMRI, T2w_sMRI, fMRI,
early-fusion concat} x 24 seeds = 480 cVAEs, hidden [110, 110], latent 10, C = 29, batch 256,
N = 1000 subjects (800 bootstrap training rows / 200 test rows per fold).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np
import pandas as pd
import torch

from . import pipeline, synthetic
from .utils import COLUMNS_NAME_AAL116

MODALITIES = ("T1w_sMRI", "T2w_sMRI", "fMRI")
EARLY = "early_fusion_modalities_HCPimage"


def train_flops_per_sample(d: int, c: int = 29, hidden: Sequence[int] = (110, 110), z: int = 10) -> int:
    """Algorithmic training FLOPs per sample per step (SURVEY 8d / BASELINE.md):
    2*[3*MAC_fwd - (D+C)*H1 - C*H_L]: forward + weight-gradient everywhere, data-gradient only
    where it is needed (not into x, not into the covariate columns of the decoder input)."""
    h = list(hidden)
    enc = [(d + c, h[0])] + [(h[i - 1], h[i]) for i in range(1, len(h))] + [(h[-1], 2 * z)]
    dec = [(z + c, h[-1])] + [(h[len(h) - i], h[len(h) - 1 - i]) for i in range(1, len(h))] + [(h[0], d)]
    mac = sum(a * b for a, b in enc + dec)
    return 2 * (3 * mac - (d + c) * h[0] - c * h[-1])


def train_param_count(d: int, c: int = 29, hidden: Sequence[int] = (110, 110), z: int = 10) -> int:
    """Trainable scalars of one single-modality cVAE (weights + biases + logvar_out); D=116: 60 092."""
    h = list(hidden)
    enc = [(d + c, h[0])] + [(h[i - 1], h[i]) for i in range(1, len(h))] + [(h[-1], z), (h[-1], z)]
    dec = [(z + c, h[-1])] + [(h[len(h) - i], h[len(h) - 1 - i]) for i in range(1, len(h))] + [(h[0], d)]
    return sum(a * b + b for a, b in enc + dec) + d


def train_bytes_per_epoch(n_rows: int, batch: int, d: int, c: int = 29, hidden: Sequence[int] = (110, 110), z: int = 10) -> int:
    """Algorithmic HBM bytes of one epoch of one model when the optimiser state is NOT resident on chip:
    per minibatch step 4 P (3 reads + 3 writes) for p, m, v (gradients stay on chip), plus every training row once,
    4 (D + C) bytes."""
    steps = -(-n_rows // batch)
    return steps * 24 * train_param_count(d, c, hidden, z) + 4 * n_rows * (d + c)


def forward_flops_per_sample(d: int, c: int = 29, hidden: Sequence[int] = (110, 110), z: int = 10) -> int:
    h = list(hidden)
    enc = [(d + c, h[0])] + [(h[i - 1], h[i]) for i in range(1, len(h))] + [(h[-1], 2 * z)]
    dec = [(z + c, h[-1])] + [(h[len(h) - i], h[len(h) - 1 - i]) for i in range(1, len(h))] + [(h[0], d)]
    return 2 * sum(a * b for a, b in enc + dec)


@dataclass
class HostWorkload:
    """Host-side arrays of one fold x modality grid (no torch, no GPU)."""
    folds: List[pipeline.FoldData]
    names: List[str]                       # modality names incl. early fusion
    dims: Dict[str, int]
    hc_label: int = 1
    hidden: Sequence[int] = (110, 110)
    latent: int = 10
    c_dim: int = 29
    batch: int = 256
    # the raw frames the folds were cut from (what pd.read_csv returns): input of the GPU prologue
    subjects: pd.DataFrame = None
    features: Dict[str, pd.DataFrame] = None
    columns: Dict[str, List[str]] = None
    n_splits: int = 5


def build_host_workload(n_subjects: int = 1000, d: int = 116, n_splits: int = 5, early_fusion: bool = True,
                        seed: int = 42, hidden=(110, 110), latent: int = 10) -> HostWorkload:
    data = synthetic.make_hcpimage(n_subjects, d, MODALITIES, seed=seed)
    subjects = data["subjects"]
    cols = COLUMNS_NAME_AAL116 if d == 116 else [f"roi_{i}" for i in range(d)]
    feats, columns = {}, {}
    parts = [pd.DataFrame({"IID": subjects["IID"]})]
    for name, x in data["features"].items():
        df = pd.DataFrame(x, columns=cols)
        df.insert(0, "IID", subjects["IID"].to_numpy())
        feats[name], columns[name] = df, list(cols)
        parts.append(pd.DataFrame(x, columns=[f"{c}_{name}" for c in cols]))   # early_fusion_modalities.py:10-35
    fused = pd.concat(parts, axis=1)
    if early_fusion:
        feats[EARLY] = fused
        columns[EARLY] = [c for c in fused.columns if c != "IID"]
    folds = pipeline.prepare_folds(subjects, feats, columns, hc_label=1, n_splits=n_splits)
    return HostWorkload(folds=folds, names=list(feats), dims={k: len(v) for k, v in columns.items()},
                        hidden=tuple(hidden), latent=latent, subjects=subjects, features=feats, columns=columns,
                        n_splits=n_splits)


def init_state_dict(d: int, hidden, latent: int, c_dim: int, seed: int):
    """Reference initialisation: ``torch.manual_seed(seed)`` then ``cVAE_multimodal([d], ...)``
    (train script :119, :161-169) -- bit-identical to the reference for the same seed."""
    from .cVAE import cVAE_multimodal
    torch.manual_seed(seed)
    m = cVAE_multimodal([d], list(hidden), latent, c_dim, learning_rate=1e-4, modalities=1, non_linear=True)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


@dataclass
class DeviceWorkload:
    specs: list
    tags: list                              # (fold, modality name, seed)
    test_xc: list                           # per member: [packed test rows]
    test_labels: list                       # per member: uint8 [N_test], 1 = patient
    train_hc_mask: list                     # per member: uint8 [N_train], 1 = healthy-control row
    host_buffers: dict = field(default_factory=dict)   # pinned (x, c) per (fold, modality) for the e2e leg
    packed: dict = field(default_factory=dict)         # (fold, modality) -> packed train tensor
    flops_per_epoch: float = 0.0
    bytes_per_epoch: float = 0.0            # algorithmic HBM bytes with streamed optimiser state (train_bytes_per_epoch)
    samples_per_epoch: int = 0
    test_subjects: int = 0


def to_device(hw: HostWorkload, device, n_seeds: int = 24, seed0: int = 0, members: Sequence[int] = None,
              pin: bool = False, gpu_prologue: bool = False) -> DeviceWorkload:
    """Upload datasets once, create one MemberSpec per (fold, modality, seed).  `members` optionally
    restricts to a subset of the global member list (multi-GPU sharding).  gpu_prologue: the packed rows are built on
    the device from the raw float64 tables (prologue.prepare_folds_gpu: RobustScaler, covariate bins, gather, packing)
    instead of being packed from the host pipeline's arrays; the results are bit-identical."""
    from .ensemble import MemberSpec, pack_rows
    gfolds = None
    if gpu_prologue:
        from . import prologue
        gfolds = prologue.prepare_folds_gpu(hw.subjects, hw.features, hw.columns, hw.hc_label, device, n_splits=hw.n_splits)
    grid = [(f, name, s) for f in range(len(hw.folds)) for name in hw.names for s in range(n_seeds)]
    if members is not None:
        grid = [grid[i] for i in members]
    wl = DeviceWorkload(specs=[], tags=[], test_xc=[], test_labels=[], train_hc_mask=[])
    tests, labels, masks, inits = {}, {}, {}, {}
    for f, name, s in grid:
        fd = hw.folds[f]
        key = (f, name)
        if key not in wl.packed:
            x = torch.from_numpy(fd.train_x[name])
            c = torch.from_numpy(fd.train_c)
            if pin:
                x, c = x.pin_memory(), c.pin_memory()
                wl.host_buffers[key] = (x, c)
            if gfolds is not None:
                wl.packed[key], tests[key] = gfolds[f].train[name], gfolds[f].test[name]
            else:
                wl.packed[key] = pack_rows(x.to(device, non_blocking=True), c.to(device, non_blocking=True))
                tests[key] = pack_rows(torch.from_numpy(fd.test_x[name]).to(device), torch.from_numpy(fd.test_c).to(device))
            labels[key] = torch.from_numpy((fd.test_df["DIA"].to_numpy() != hw.hc_label).astype(np.uint8)).to(device)
            masks[key] = torch.from_numpy((fd.train_df["DIA"].to_numpy() == hw.hc_label).astype(np.uint8)).to(device)
        d = hw.dims[name]
        if (d, s) not in inits:
            inits[(d, s)] = init_state_dict(d, hw.hidden, hw.latent, hw.c_dim, 42 + seed0 + s)
        wl.specs.append(MemberSpec([d], list(hw.hidden), hw.latent, hw.c_dim, [wl.packed[key]], batch=hw.batch,
                                   seed=1000003 * (seed0 + s) + 7919 * f + hw.names.index(name),
                                   state_dict=inits[(d, s)], tag=(f, name, seed0 + s)))
        wl.tags.append((f, name, seed0 + s))
        wl.test_xc.append([tests[key]]); wl.test_labels.append(labels[key]); wl.train_hc_mask.append(masks[key])
        n_tr = fd.train_x[name].shape[0]
        wl.samples_per_epoch += n_tr
        wl.flops_per_epoch += n_tr * train_flops_per_sample(d, hw.c_dim, hw.hidden, hw.latent)
        wl.bytes_per_epoch += train_bytes_per_epoch(n_tr, hw.batch, d, hw.c_dim, hw.hidden, hw.latent)
        wl.test_subjects += fd.test_x[name].shape[0]
    return wl
