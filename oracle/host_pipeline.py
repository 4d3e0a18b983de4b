"""numpy restatement of the reference's host-side fold / loading / covariate path.

TEST INFRASTRUCTURE ONLY.  Integer outputs here must be bit-exact with the reference.

Restated behaviour:
* ``generate_kfold_ids`` (KFold over HC+other, bootstrap through the GLOBAL numpy RNG)
                                                   utils.py:73-93
* ``generate_kfold_ids_with_unigroup``             utils.py:50-70
* ``load_dataset`` / ``load_demographic_data``     utils.py:112-168 (merge row order)
* RobustScaler fit/transform                       ..._train_cvae_supervised.py:101-102
* rank(method='first') -> qcut -> one-hot          ..._train_cvae_supervised.py:107-114
* ``MyDataset_labels`` int64 cast                  utils_vae.py:24
sklearn's KFold(shuffle=True, random_state=42) is restated from its published
algorithm (scikit-learn 1.6.1 pinned in environment.yml:327; identical in 1.9.0).
"""
from __future__ import annotations

import numpy as np


def kfold_indices(n: int, n_splits: int, seed: int = 42):
    """sklearn ``KFold(n_splits, shuffle=True, random_state=seed).split(range(n))``.

    idx = arange(n); RandomState(seed).shuffle(idx); contiguous fold slices of size
    n//k (+1 for the first n%k); test = idx[slice] sorted (sklearn builds a boolean
    mask), train = sorted complement.
    """
    idx = np.arange(n)
    np.random.RandomState(seed).shuffle(idx)
    sizes = np.full(n_splits, n // n_splits, dtype=int)
    sizes[: n % n_splits] += 1
    out, start = [], 0
    for s in sizes:
        mask = np.zeros(n, dtype=bool)
        mask[idx[start:start + s]] = True
        out.append((np.nonzero(~mask)[0], np.nonzero(mask)[0]))
        start += s
    return out


def bootstrap_folds(n: int, n_splits: int, oversample: float = 1.0, seed: int = 42):
    """Row positions (into the HC+other concatenation) of utils.py:73-93.

    ``np.random.seed(42)`` in main (train script :43) then, per fold in order,
    ``np.random.choice(train_ids, size=int(len*O), replace=True)`` == indexing with
    ``RandomState.randint(0, len, size)`` on the same legacy stream.
    Returns [(boot_train_positions, test_positions)] per fold.
    """
    rs = np.random.RandomState(seed)
    out = []
    for tr, te in kfold_indices(n, n_splits, 42):
        size = int(len(tr) * oversample)
        pick = rs.randint(0, len(tr), size=size)
        out.append((tr[pick], te))
    return out


def merge_rows(feature_ids, selected_ids):
    """Row order of ``pd.merge(features, merge(ids, demo))`` (utils.py:118-120, 158).

    Inner merge keeps the LEFT (feature file) order; each feature row is repeated once
    per occurrence of its id among ``selected_ids`` (bootstrap duplicates stay adjacent).
    Returns positions into the feature file.
    """
    counts = {}
    for s in selected_ids:
        counts[s] = counts.get(s, 0) + 1
    rows = []
    for pos, fid in enumerate(feature_ids):
        rows.extend([pos] * counts.get(fid, 0))
    return np.asarray(rows, dtype=np.int64)


def robust_scale_fit(x):
    """sklearn RobustScaler: center = median, scale = q75 - q25 (linear interp); 0 -> 1."""
    x = np.asarray(x, dtype=np.float64)
    center = np.nanmedian(x, axis=0)
    q = np.nanpercentile(x, (25.0, 75.0), axis=0)
    scale = q[1] - q[0]
    scale[scale < 10 * np.finfo(scale.dtype).eps] = 1.0
    return center, scale


def robust_scale_apply(x, center, scale):
    return (np.asarray(x, dtype=np.float64) - center) / scale


def rank_first(v):
    """pandas ``Series.rank(method='first')``: 1-based rank, ties by order of appearance."""
    order = np.argsort(np.asarray(v), kind="stable")
    r = np.empty(len(v), dtype=np.float64)
    r[order] = np.arange(1, len(v) + 1)
    return r


def qcut_rank_bins(v, q, pandas_ge_22=True):
    """``pd.qcut(v.rank(method='first'), q, labels=range(q))`` via pandas' own algorithm:
    edges = ``np.quantile`` (linear) of the ranks at linspace(0,1,q+1); bins are
    right-closed with the lowest edge included (``searchsorted(side='left') - 1``).

    Version drift: the reference pins pandas 2.0.3 (environment.yml:307),
    which uses the plain linspace.  pandas >= 2.2 (3.0.2 in this image, i.e. what the
    reference does when it is run here) first rounds every quantile that is not exactly
    representable *up* by one ulp (``np.nextafter``).  The two differ for some n
    (e.g. 1000, 37); ``pandas_ge_22`` selects which to restate.  The golden vectors were
    produced with pandas 3.0.2.
    """
    r = rank_first(v)
    qs = np.linspace(0, 1, q + 1)
    if pandas_ge_22:
        np.putmask(qs, q * qs != np.arange(q + 1), np.nextafter(qs, 1))
    edges = np.quantile(r, qs)
    b = np.searchsorted(edges, r, side="left") - 1
    b[r == edges[0]] = 0
    return b.astype(np.int64)


def covariate_onehots(age, sex, n_age=27, n_sex=2):
    """[N, 29] float32 one-hots (train script :105-126), cast to int64 by the dataset."""
    a = np.eye(n_age)[qcut_rank_bins(age, n_age)]
    s = np.eye(n_sex)[qcut_rank_bins(sex, n_sex)]
    return np.concatenate((a, s), axis=1).astype("float32")
