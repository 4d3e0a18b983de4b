"""Deviation scoring on the GPU: per-ROI / per-subject deviations, healthy-control-referenced
z-scores and ROC-AUC, batched over (member, modality) segments.

Replaces multimodal_kfold_test_cvae_supervised.py:112-141 (deviations),
utils_vae.py:147-161 (deviation helpers) and
multimodal_kfold_cvae_group_analysis_1x1.py:123-124 (roc_curve + auc).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from .ensemble import _stream_ptr


def _dev(ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError("scoring needs CUDA tensors (libnmb has no CPU fallback)")
            return t.device
    raise ValueError("no tensors")


def _seg_tables(x, xhat):
    n_rows, d, ldx = [], [], []
    for a, b in zip(x, xhat):
        if a.dtype != torch.float32 or b.dtype != torch.float32 or not a.is_contiguous() or not b.is_contiguous():
            raise ValueError("float32 contiguous tensors required")
        if a.shape[0] != b.shape[0] or a.shape[1] < b.shape[1]:
            raise ValueError("x rows must be packed rows whose first D columns are the ROIs of xhat")
        n_rows.append(a.shape[0]); d.append(b.shape[1]); ldx.append(a.shape[1])
    return n_rows, d, ldx


def normative_stats(x: Sequence[torch.Tensor], xhat: Sequence[torch.Tensor],
                    mask: Optional[Sequence[Optional[torch.Tensor]]] = None) -> List[torch.Tensor]:
    """Per-ROI (mean, population std) of (x - xhat)^2 over the reference rows; returns [2, D] each."""
    dev = _dev(x)
    n_rows, d, ldx = _seg_tables(x, xhat)
    out = [torch.empty((2, dd), dtype=torch.float32, device=dev) for dd in d]
    mk = None
    if mask is not None:
        mk = [None if m is None else m.to(device=dev, dtype=torch.uint8).contiguous() for m in mask]
    with torch.cuda.device(dev):
        _lib.check(_lib.load().nmb_normative_stats(
            len(x), _lib.ptr_table([t.data_ptr() for t in x]), _lib.int_table(ldx),
            _lib.ptr_table([t.data_ptr() for t in xhat]),
            _lib.ptr_table([None if m is None else m.data_ptr() for m in mk]) if mk is not None else None,
            _lib.int_table(n_rows), _lib.int_table(d), _lib.ptr_table([t.data_ptr() for t in out]),
            _stream_ptr(dev)), "nmb_normative_stats")
    return out


def deviation(x: Sequence[torch.Tensor], xhat: Sequence[torch.Tensor],
              stats: Optional[Sequence[torch.Tensor]] = None, want_roi: bool = True):
    """Returns (dev_roi[s] [N,D] or None, z[s] [N,D] or None, dev_subj[s] [N])."""
    dev = _dev(x)
    n_rows, d, ldx = _seg_tables(x, xhat)
    roi = [torch.empty((n, dd), dtype=torch.float32, device=dev) for n, dd in zip(n_rows, d)] if want_roi else None
    z = [torch.empty((n, dd), dtype=torch.float32, device=dev) for n, dd in zip(n_rows, d)] if stats is not None else None
    subj = [torch.empty((n,), dtype=torch.float32, device=dev) for n in n_rows]
    with torch.cuda.device(dev):
        _lib.check(_lib.load().nmb_deviation(
            len(x), _lib.ptr_table([t.data_ptr() for t in x]), _lib.int_table(ldx),
            _lib.ptr_table([t.data_ptr() for t in xhat]),
            _lib.ptr_table([t.data_ptr() for t in stats]) if stats is not None else None,
            _lib.int_table(n_rows), _lib.int_table(d),
            _lib.ptr_table([t.data_ptr() for t in roi]) if roi is not None else None,
            _lib.ptr_table([t.data_ptr() for t in z]) if z is not None else None,
            _lib.ptr_table([t.data_ptr() for t in subj]), _stream_ptr(dev)), "nmb_deviation")
    return roi, z, subj


def auc(scores: Sequence[torch.Tensor], labels: Sequence[torch.Tensor], want_pairs: bool = False):
    """ROC-AUC of every column of scores[s] ([N, K] or [N]) vs labels[s] (1 = patient).

    Exact pair counting with tie half-credit (== sklearn roc_curve + auc).  Returns float64
    tensors [K] (and the uint64 pair counts U2 when want_pairs)."""
    dev = _dev(scores)
    sc = [s.reshape(s.shape[0], -1).to(torch.float32).contiguous() for s in scores]
    lb = [l.to(device=dev, dtype=torch.uint8).contiguous() for l in labels]
    n_rows = [s.shape[0] for s in sc]
    n_cols = [s.shape[1] for s in sc]
    out = [torch.empty((k,), dtype=torch.float64, device=dev) for k in n_cols]
    u2 = [torch.zeros((k,), dtype=torch.int64, device=dev) for k in n_cols] if want_pairs else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().nmb_auc(
            len(sc), _lib.ptr_table([t.data_ptr() for t in sc]), _lib.ptr_table([t.data_ptr() for t in lb]),
            _lib.int_table(n_rows), _lib.int_table(n_cols), _lib.ptr_table([t.data_ptr() for t in out]),
            _lib.ptr_table([t.data_ptr() for t in u2]) if u2 is not None else None, _stream_ptr(dev)), "nmb_auc")
    return (out, u2) if want_pairs else out


def mean_rows(vectors: Sequence[torch.Tensor]) -> torch.Tensor:
    """Elementwise mean of k equally-shaped vectors (modality averaging, group analysis :212-215)."""
    dev = _dev(vectors)
    vs = [v.to(torch.float32).contiguous() for v in vectors]
    out = torch.empty_like(vs[0])
    with torch.cuda.device(dev):
        _lib.check(_lib.load().nmb_mean_rows(_lib.ptr_table([v.data_ptr() for v in vs]), len(vs), vs[0].numel(),
                                             out.data_ptr(), _stream_ptr(dev)), "nmb_mean_rows")
    return out


def philox_normal(seed: int, step: int, n: int, stream_id: int = 0, device="cuda") -> torch.Tensor:
    """The in-kernel eps stream (definition: oracle/philox.py) for inspection and tests."""
    out = torch.empty((n,), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().nmb_philox_normal(int(seed), int(step), int(stream_id), n, out.data_ptr(),
                                                 _stream_ptr(out.device)), "nmb_philox_normal")
    return out
