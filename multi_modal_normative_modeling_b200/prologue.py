"""GPU-resident preprocessing prologue: RobustScaler, rank -> quantile-bin one-hots, bootstrap / merge
gather and row packing of a fold, on the raw float64 feature table kept in HBM -- the work of
multimodal_kfold_train_cvae_supervised.py:88-131 and ..._test_cvae_supervised.py:74-102 after the CSV files have been
read.  Integer work that defines WHICH rows a fold holds (KFold, the bootstrap on the legacy numpy RNG stream, the
feature-file merge order) stays on the host and is handed over as row positions; the qcut bin edges, which depend only
on the number of rows, are pandas' own (computed once per distinct n).  Everything else is three libnmb kernels per
(fold, modality); results are bit-identical to the pandas / sklearn path of pipeline.prepare_folds."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np
import pandas as pd
import torch

from . import _lib
from .ensemble import _stream_ptr
from .pipeline import N_AGE_BINS, N_SEX_BINS, kfold_ids

_EDGES: Dict[tuple, np.ndarray] = {}


def qcut_edges(n: int, q: int) -> np.ndarray:
    """Bin edges of ``pd.qcut(ranks, q)`` for n ranks 1..n (they do not depend on the order of the ranks): pandas' own
    values, so version-specific details of its quantile arithmetic are inherited, not restated."""
    key = (n, q)
    if key not in _EDGES:
        _, edges = pd.qcut(pd.Series(np.arange(1, n + 1, dtype=np.float64)), q=q, labels=False, retbins=True)
        _EDGES[key] = np.asarray(edges, dtype=np.float64)
    return _EDGES[key]


def row_positions(feature_ids: Sequence[str], demo_ids: Sequence[str], selected_ids: Sequence[str]) -> np.ndarray:
    """Row positions (into the feature file) of ``merge(features, merge(ids, demographics))`` (utils.py:112-168):
    feature-file order, one copy per occurrence of the id among the selected (bootstrap) ids, ids without demographics
    dropped."""
    have = set(demo_ids)
    counts = pd.Series([i for i in selected_ids if i in have]).value_counts()
    rep = pd.Series(list(feature_ids)).map(counts).fillna(0).to_numpy().astype(np.int64)
    return np.repeat(np.arange(len(rep), dtype=np.int64), rep)


@dataclass
class RawModality:
    x: torch.Tensor                  # [n_subjects, D] float64 on the device, feature-file row order
    age: torch.Tensor                # [n_subjects] float64 (demographics aligned to the feature rows)
    sex: torch.Tensor
    d: int


@dataclass
class GpuFold:
    train: Dict[str, torch.Tensor] = field(default_factory=dict)      # packed fp32 rows per modality
    test: Dict[str, torch.Tensor] = field(default_factory=dict)
    center: Dict[str, torch.Tensor] = field(default_factory=dict)     # float64 [D]
    scale: Dict[str, torch.Tensor] = field(default_factory=dict)
    train_pos: np.ndarray = None
    test_pos: np.ndarray = None


def _dev_i32(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def _ptr_n(idx, n_all):
    return (idx.data_ptr(), idx.numel()) if idx is not None else (None, n_all)


def robust_fit(x: torch.Tensor, idx: torch.Tensor = None):
    """(center, scale) float64 [D] of rows idx of x (all rows if None) -- RobustScaler().fit (nmb_robust_fit)."""
    d = x.shape[1]
    center = torch.empty(d, dtype=torch.float64, device=x.device)
    scale = torch.empty(d, dtype=torch.float64, device=x.device)
    ip, n = _ptr_n(idx, x.shape[0])
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().nmb_robust_fit(x.data_ptr(), x.stride(0), d, ip, n, center.data_ptr(),
                                              scale.data_ptr(), _stream_ptr(x.device)), "nmb_robust_fit")
    return center, scale


def rank_bins(v: torch.Tensor, idx: torch.Tensor, q: int) -> torch.Tensor:
    """``pd.qcut(v[idx].rank(method='first'), q, labels=False)`` as int32 [len(idx)] (nmb_rank_bins)."""
    ip, n = _ptr_n(idx, v.shape[0])
    edges = torch.from_numpy(qcut_edges(n, q)).to(v.device)
    bins = torch.empty(n, dtype=torch.int32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.load().nmb_rank_bins(v.data_ptr(), ip, n, edges.data_ptr(), q, bins.data_ptr(),
                                             _stream_ptr(v.device)), "nmb_rank_bins")
    return bins


def pack_scaled(x, idx, center, scale, age_bin, sex_bin) -> torch.Tensor:
    d = x.shape[1]
    ip, n = _ptr_n(idx, x.shape[0])
    out = torch.empty((n, _lib.packed_row_stride(d, N_AGE_BINS + N_SEX_BINS)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().nmb_pack_rows_scaled(x.data_ptr(), x.stride(0), d, ip, n, center.data_ptr(),
                                                    scale.data_ptr(), age_bin.data_ptr(), N_AGE_BINS, sex_bin.data_ptr(),
                                                    N_SEX_BINS, out.data_ptr(), _stream_ptr(x.device)), "nmb_pack_rows_scaled")
    return out


MAX_ROWS = 8192       # rows per fold the shared-memory sort of nmb_robust_fit / nmb_rank_bins holds


def frame_to_packed(train_df: pd.DataFrame, test_df, cols, device, center_scale=None):
    """The CLI flavour: the fold's frames come out of ``load_dataset`` (CSV + merges stay pandas: the file contract);
    everything after that -- RobustScaler fit on train / transform, covariate bins from each frame's own ranks,
    packing -- runs on the device.  Returns (packed_train, packed_test or None, scaled_test_float64 or None,
    (center, scale)).  test_df may be None."""
    dev = torch.device(device)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

    def bins(df):
        return (rank_bins(up(df["AGE"].to_numpy()), None, N_AGE_BINS), rank_bins(up(df["PTGENDER"].to_numpy()), None, N_SEX_BINS))
    xtr = up(train_df[cols].to_numpy())
    center, scale = center_scale or robust_fit(xtr)
    ptr = pack_scaled(xtr, None, center, scale, *bins(train_df))
    if test_df is None:
        return ptr, None, None, (center, scale)
    xte = up(test_df[cols].to_numpy())
    pte = pack_scaled(xte, None, center, scale, *bins(test_df))
    return ptr, pte, ((xte - center) / scale), (center, scale)


def upload_raw(subjects: pd.DataFrame, features: Dict[str, pd.DataFrame], columns: Dict[str, List[str]], device):
    """Raw float64 tables to HBM, once per run (what ``pd.read_csv`` returned; no scaling, no selection)."""
    demo = subjects.dropna().set_index("IID")
    raw = {}
    for name, feat in features.items():
        x = torch.from_numpy(np.ascontiguousarray(feat[columns[name]].to_numpy(np.float64))).to(device)
        age = torch.from_numpy(feat["IID"].map(demo["AGE"]).to_numpy(np.float64)).to(device)
        sex = torch.from_numpy(feat["IID"].map(demo["PTGENDER"]).to_numpy(np.float64)).to(device)
        raw[name] = RawModality(x=x, age=age, sex=sex, d=len(columns[name]))
    return raw


def prepare_folds_gpu(subjects: pd.DataFrame, features: Dict[str, pd.DataFrame], columns: Dict[str, List[str]],
                      hc_label: int, device, n_splits: int = 5, oversample_percentage: float = 1.0,
                      training_class: str = "nm", seed: int = 42, raw=None) -> List[GpuFold]:
    """The GPU counterpart of pipeline.prepare_folds: same folds, same bootstrap stream, packed rows built on the device."""
    if not torch.cuda.is_available():
        raise RuntimeError("the GPU prologue needs a CUDA device (libnmb has no CPU fallback)")
    device = torch.device(device)
    raw = raw or upload_raw(subjects, features, columns, device)
    np.random.seed(seed)                                  # train script :43
    demo_ids = subjects.dropna()["IID"].to_numpy()
    folds = []
    for train_ids, test_ids in kfold_ids(subjects, hc_label, n_splits, oversample_percentage, training_class):
        gf = GpuFold()
        for name, feat in features.items():
            r = raw[name]
            fid = feat["IID"].to_numpy()
            tr = _dev_i32(row_positions(fid, demo_ids, train_ids), device)
            te = _dev_i32(row_positions(fid, demo_ids, test_ids), device)
            center, scale = robust_fit(r.x, tr)           # fitted on TRAIN, applied to both (test script :83-90)
            gf.center[name], gf.scale[name] = center, scale
            # covariate bins from each set's own ranks (train script :105-114, test script :93-97)
            gf.train[name] = pack_scaled(r.x, tr, center, scale, rank_bins(r.age, tr, N_AGE_BINS), rank_bins(r.sex, tr, N_SEX_BINS))
            gf.test[name] = pack_scaled(r.x, te, center, scale, rank_bins(r.age, te, N_AGE_BINS), rank_bins(r.sex, te, N_SEX_BINS))
            gf.train_pos, gf.test_pos = tr, te
        folds.append(gf)
    return folds
