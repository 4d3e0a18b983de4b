"""Per-fold data preparation shared by the CLIs and the benchmark: the host half of
multimodal_kfold_train_cvae_supervised.py:82-131 and
multimodal_kfold_test_cvae_supervised.py:74-102, factored so that all folds / modalities of a
run are prepared once and handed to the fused ensemble kernels.

Everything here is pandas / sklearn on the host *by design* (bit-exact indexing); the calls are
the ones the reference makes: ``pd.merge`` (row order), ``RobustScaler``,
``Series.rank(method='first')`` + ``pd.qcut`` (27 age bins + 2 "gender" bins).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np
import pandas as pd
from sklearn.model_selection import KFold
from sklearn.preprocessing import RobustScaler

N_AGE_BINS, N_SEX_BINS = 27, 2


def covariate_onehots(df: pd.DataFrame) -> np.ndarray:
    """[N, 29] float32: rank -> quantile bin -> one-hot for AGE (27) and PTGENDER (2)
    (train script :105-126; the "gender" bins are a rank-based 50/50 split)."""
    age_bins = pd.qcut(df["AGE"].rank(method="first"), q=N_AGE_BINS, labels=list(range(N_AGE_BINS)))
    sex_bins = pd.qcut(df["PTGENDER"].rank(method="first"), q=N_SEX_BINS, labels=list(range(N_SEX_BINS)))
    one_hot_age = np.eye(N_AGE_BINS)[np.asarray(age_bins.values, dtype=np.int64)]
    one_hot_sex = np.eye(N_SEX_BINS)[np.asarray(sex_bins.values, dtype=np.int64)]
    return np.concatenate((one_hot_age, one_hot_sex), axis=1).astype("float32")


def select_rows(features: pd.DataFrame, demographics: pd.DataFrame, ids: Sequence[str]) -> pd.DataFrame:
    """In-memory ``load_dataset`` (utils.py:112-168): merge(ids, demographics) then
    merge(features, .) on IID -- feature-file row order, bootstrap duplicates adjacent."""
    ids_df = pd.DataFrame({"IID": np.asarray(ids)})
    ids_df["participant_id"] = ids_df["IID"]
    demo = pd.merge(ids_df, demographics.dropna(), on="IID")
    return pd.merge(features, demo, on="IID")


def kfold_ids(subjects: pd.DataFrame, hc_label: int, n_splits: int, oversample_percentage: float = 1.0,
              training_class: str = "nm", disease_label: int = 0):
    """In-memory ``generate_kfold_ids`` (utils.py:73-93) called as in the train script :53-66.
    Consumes the GLOBAL numpy RNG exactly like the reference (caller seeds it with 42)."""
    label = hc_label if training_class == "nm" else disease_label
    group = pd.concat([subjects[subjects["DIA"] == label], subjects[subjects["DIA"] != label]])
    out = []
    for train_idx, test_idx in KFold(n_splits=n_splits, shuffle=True, random_state=42).split(group):
        train_ids = group.iloc[train_idx]["IID"]
        boot = np.random.choice(train_ids, size=int(len(train_ids) * oversample_percentage), replace=True)
        out.append((np.asarray(boot), group.iloc[test_idx]["IID"].to_numpy()))
    return out


@dataclass
class FoldData:
    fold: int
    names: List[str]
    train_x: Dict[str, np.ndarray] = field(default_factory=dict)   # RobustScaler-ed, float32 [Ntr, D]
    test_x: Dict[str, np.ndarray] = field(default_factory=dict)    # scaler fitted on train (test script :83-90)
    test_x64: Dict[str, np.ndarray] = field(default_factory=dict)  # the float64 frame the reference subtracts from
    train_c: np.ndarray = None                                     # [Ntr, 29] float32 one-hots
    test_c: np.ndarray = None                                      # from the TEST set's own ranks (test script :93-97)
    train_df: pd.DataFrame = None
    test_df: pd.DataFrame = None


def prepare_folds(subjects: pd.DataFrame, features: Dict[str, pd.DataFrame], columns: Dict[str, List[str]],
                  hc_label: int, n_splits: int = 5, oversample_percentage: float = 1.0,
                  training_class: str = "nm", hc_only: bool = False, seed: int = 42) -> List[FoldData]:
    """All folds x modalities of one run.  ``hc_only`` = the nmmlp variant, which keeps only the
    healthy-control training rows (multimodal_kfold_cvae_nmmlp.py:314)."""
    np.random.seed(seed)                                  # train script :43
    folds = []
    for fold, (train_ids, test_ids) in enumerate(
            kfold_ids(subjects, hc_label, n_splits, oversample_percentage, training_class)):
        fd = FoldData(fold=fold, names=list(features))
        for name, feat in features.items():
            tr = select_rows(feat, subjects, train_ids)
            te = select_rows(feat, subjects, test_ids)
            if hc_only:
                tr = tr.loc[tr["DIA"] == hc_label]
            scaler = RobustScaler()
            fd.train_x[name] = scaler.fit_transform(tr[columns[name]].values).astype(np.float32)
            te64 = scaler.transform(te[columns[name]].values)
            fd.test_x64[name] = te64
            fd.test_x[name] = te64.astype(np.float32)
            fd.train_c = covariate_onehots(tr)
            fd.test_c = covariate_onehots(te)             # last modality wins, like the reference (:102)
            fd.train_df, fd.test_df = tr, te
        folds.append(fd)
    return folds
