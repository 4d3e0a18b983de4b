"""numpy restatement of the reference's deviation scoring and group analysis.

TEST INFRASTRUCTURE ONLY.

Restated behaviour:
* per-subject deviation  sum_d (x - xhat)^2 / D         cVAE.py:1210-1211, utils_vae.py:147-148
* per-ROI deviation      (x - xhat)^2                   ..._test_cvae_supervised.py:141, utils_vae.py:151-152
* latent z-score         (mu - mean(mu_tr)) / sqrt(var(mu_tr) + var_s)   utils_vae.py:155-161
* ROC-AUC / Youden / acc / sens / spec                  ..._group_analysis_1x1.py:105-157
  (sklearn ``roc_curve`` + ``auc`` restated as the Mann-Whitney statistic with tie
  half-credit; equality to 1e-16 is pinned by tests/golden/host_callsites.npz)

The reference has NO ROI-space z-score (SURVEY.md section 8a note).  The framework
defines it here, once, and the CUDA kernels follow this definition:

    r_id  = (x_id - xhat_id)^2                       (squared residual, a15)
    m_d   = mean_{i in HC reference rows} r_id
    s_d   = std_{i in HC reference rows}  r_id       (ddof = 0, as np.var in utils_vae.py:156)
    z_id  = (r_id - m_d) / s_d
"""
from __future__ import annotations

import numpy as np


def recon_deviation(x, x_pred):
    x = np.asarray(x, dtype=np.float64)
    return np.sum((x - x_pred) ** 2, axis=1) / x.shape[1]


def recon_deviation_roi(x, x_pred):
    return (np.asarray(x, dtype=np.float64) - x_pred) ** 2


def normative_stats(r_ref):
    """Per-ROI mean and population std over the HC reference rows.  r_ref: [N_ref, D]."""
    r_ref = np.asarray(r_ref, dtype=np.float64)
    return r_ref.mean(axis=0), r_ref.std(axis=0)


def zscores(r, mean, std):
    return (np.asarray(r, dtype=np.float64) - mean) / std


def latent_zscores(mu_train, mu_sample, var_sample):
    var = np.var(mu_train, axis=0)
    return (mu_sample - np.mean(mu_train, axis=0)) / np.sqrt(var + var_sample)


def latent_deviation(mu_train, mu_sample, var_sample):
    return np.sum(np.abs(latent_zscores(mu_train, mu_sample, var_sample)), axis=1) / mu_sample.shape[1]


def auc_pairs(scores, labels):
    """Exact integer pair count U2 = sum_{pos,neg} (2*[s_p > s_n] + [s_p == s_n]).

    AUC = U2 / (2 * n_pos * n_neg).  This is what the CUDA AUC kernel counts, so the
    comparison with it is bit-exact.
    """
    scores = np.asarray(scores)
    labels = np.asarray(labels)
    pos = scores[labels == 1]
    neg = np.sort(scores[labels == 0])
    less = np.searchsorted(neg, pos, side="left")
    leq = np.searchsorted(neg, pos, side="right")
    return int(np.sum(2 * less + (leq - less))), len(pos), len(neg)


def auc(scores, labels):
    """== sklearn ``auc(*roc_curve(labels, scores)[:2])`` (rank-sum with average ranks)."""
    u2, n1, n0 = auc_pairs(scores, labels)
    if n1 == 0 or n0 == 0:
        return float("nan")
    return u2 / (2.0 * n1 * n0)


def classification_performance(scores, labels):
    """``compute_classification_performance`` method='roc', training_class='nm'.

    labels: 0 = HC, 1 = patient.  Youden's J over the ROC operating points in
    sklearn's order (thresholds descending, first maximum wins; the sentinel first
    point has J = 0).  Returns (auc, acc, sens, spec, sig_ratio, threshold).
    """
    scores = np.asarray(scores, dtype=np.float64)
    labels = np.asarray(labels)
    order = np.argsort(-scores, kind="mergesort")
    s, y = scores[order], labels[order]
    last_of_run = np.r_[np.nonzero(np.diff(s))[0], len(s) - 1]
    tps = np.cumsum(y == 1)[last_of_run]
    fps = np.cumsum(y == 0)[last_of_run]
    n1, n0 = int((labels == 1).sum()), int((labels == 0).sum())
    j = tps / n1 - fps / n0
    best = int(np.argmax(j))
    thr = s[last_of_run][best] if j[best] > 0 else np.inf
    pred = (scores >= thr).astype(int)
    acc = float((pred == labels).mean())
    tp = np.sum((pred == 1) & (labels == 1)); fn = np.sum((pred == 0) & (labels == 1))
    tn = np.sum((pred == 0) & (labels == 0)); fp = np.sum((pred == 1) & (labels == 0))
    a = auc(scores, labels)
    return a, acc, tp / (tp + fn), tn / (tn + fp), a / (1 - a), float(thr)
