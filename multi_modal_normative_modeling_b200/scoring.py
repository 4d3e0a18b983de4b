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


def latent_deviation(mu_train: Sequence[torch.Tensor], mu: Sequence[torch.Tensor], logvar: Sequence[torch.Tensor],
                     want_separate: bool = True):
    """The reference's latent-space normative deviation (utils_vae.py:155-161) per segment:
    ``separate_latent_deviation`` z[s] [N, Z] and ``latent_deviation`` dev[s] [N].
    mu_train[s]: latent means of the reference (healthy-control training) rows; mu / logvar[s]: the scored rows
    (var_sample = exp(logvar), what ``pred_latent`` returns)."""
    dev = _dev(mu)
    f = lambda ts: [t.to(device=dev, dtype=torch.float32).contiguous() for t in ts]
    mt, m, lv = f(mu_train), f(mu), f(logvar)
    for a, b, c in zip(mt, m, lv):
        if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1] or b.shape != c.shape:
            raise ValueError("mu_train [Nt, Z], mu [N, Z], logvar [N, Z] required")
    z = [torch.empty_like(t) for t in m] if want_separate else None
    out = [torch.empty((t.shape[0],), dtype=torch.float32, device=dev) for t in m]
    with torch.cuda.device(dev):
        _lib.check(_lib.load().nmb_latent_deviation(
            len(m), _lib.ptr_table([t.data_ptr() for t in mt]), _lib.int_table([t.shape[0] for t in mt]),
            _lib.ptr_table([t.data_ptr() for t in m]), _lib.ptr_table([t.data_ptr() for t in lv]),
            _lib.int_table([t.shape[0] for t in m]), _lib.int_table([t.shape[1] for t in m]),
            _lib.ptr_table([t.data_ptr() for t in z]) if z is not None else None,
            _lib.ptr_table([t.data_ptr() for t in out]), _stream_ptr(dev)), "nmb_latent_deviation")
    return z, out


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


class DeviationScorer:
    """Whole-ensemble deviation scoring with all buffers and argument tables prepared ONCE.

    One ``run()`` = the test script + group analysis of the reference for every member
    (multimodal_kfold_test_cvae_supervised.py:107-141, ..._group_analysis_1x1.py:105-157):
    reconstruct the training and the test rows, per-ROI normative statistics over the healthy-control
    training rows, per-ROI / per-subject deviations and z-scores of the test rows, ROC-AUC of every
    ROI z-score and of the subject score -- six libnmb launches, no per-member Python work.
    Segments are (member, modality) pairs in member-major order."""

    def __init__(self, trainer, train_xc, test_xc, train_hc_mask, test_labels, mode: str = "mean",
                 params_untouched: bool = False):
        """params_untouched: the caller promises not to write the trainer's packed parameter tensors between training
        and scoring, so the weight planes the training kernel left behind are reused (NMB_RECON_KEEP_PLANES) instead of
        being rebuilt for each of the two reconstruction launches."""
        self.tr = trainer
        self.lib = _lib.load()
        self.dev = trainer.device
        self.mode = (_lib.RECON_MEAN if mode == "mean" else _lib.RECON_SAMPLE) | (_lib.RECON_KEEP_PLANES if params_untouched else 0)
        n = trainer.n
        dev = self.dev
        self._keep = [train_xc, test_xc, train_hc_mask, test_labels]
        seg_member, seg_d, seg_ldx, n_tr, n_te = [], [], [], [], []
        for i, s in enumerate(trainer.specs):
            for k, d in enumerate(s.input_dims):
                seg_member.append(i); seg_d.append(int(d)); seg_ldx.append(int(train_xc[i][k].shape[1]))
                n_tr.append(int(train_xc[i][k].shape[0])); n_te.append(int(test_xc[i][k].shape[0]))
        self.seg_member, self.seg_d, self.n_seg = seg_member, seg_d, len(seg_d)
        self.n_train, self.n_test = n_tr, n_te
        import numpy as np
        d = np.asarray(seg_d, np.int64); ntr = np.asarray(n_tr, np.int64); nte = np.asarray(n_te, np.int64)
        off = lambda sizes: np.concatenate([[0], np.cumsum((sizes + 3) // 4 * 4)])     # 16-byte aligned segments
        self.o_hat_tr, self.o_hat_te, self.o_stats = off(ntr * d), off(nte * d), off(2 * d)
        self.o_subj, self.o_auc = off(nte), off(d)
        f32 = lambda m: torch.empty((int(m),), dtype=torch.float32, device=dev)
        self.xhat_train, self.xhat_test = f32(self.o_hat_tr[-1]), f32(self.o_hat_te[-1])
        self.stats, self.roi, self.z, self.subj = f32(self.o_stats[-1]), f32(self.o_hat_te[-1]), f32(self.o_hat_te[-1]), f32(self.o_subj[-1])
        self.auc_roi = torch.empty((int(self.o_auc[-1]),), dtype=torch.float64, device=dev)
        self.auc_subj = torch.empty((self.n_seg,), dtype=torch.float64, device=dev)
        masks = [None if m is None else m.to(device=dev, dtype=torch.uint8).contiguous() for m in train_hc_mask]
        labels = [l.to(device=dev, dtype=torch.uint8).contiguous() for l in test_labels]
        self._keep += [masks, labels]
        M = _lib.NMB_MAX_MOD

        def member_table(xc, base, offsets):
            xt, ot = [None] * (n * M), [None] * (n * M)
            s = 0
            for i, sp in enumerate(trainer.specs):
                for k in range(len(sp.input_dims)):
                    xt[i * M + k] = xc[i][k].data_ptr(); ot[i * M + k] = base.data_ptr() + 4 * int(offsets[s]); s += 1
            return _lib.ptr_table(xt), _lib.ptr_table(ot)
        self.t_xc_tr, self.t_hat_tr = member_table(train_xc, self.xhat_train, self.o_hat_tr)
        self.t_xc_te, self.t_hat_te = member_table(test_xc, self.xhat_test, self.o_hat_te)
        self.t_rows_tr = _lib.int_table([int(train_xc[i][0].shape[0]) for i in range(n)])
        self.t_rows_te = _lib.int_table([int(test_xc[i][0].shape[0]) for i in range(n)])
        # both row sets in one launch (nmb_ensemble_reconstruct_sets): tables of set 0 (training rows) then set 1 (test rows)
        cat = lambda a, b: _lib.ptr_table([a[i] for i in range(n * M)] + [b[i] for i in range(n * M)])
        self.t_xc_both, self.t_hat_both = cat(self.t_xc_tr, self.t_xc_te), cat(self.t_hat_tr, self.t_hat_te)
        self.t_rows_both = _lib.int_table([int(train_xc[i][0].shape[0]) for i in range(n)] + [int(test_xc[i][0].shape[0]) for i in range(n)])
        seg = range(self.n_seg)
        ptrs = lambda base, offsets, size: _lib.ptr_table([base.data_ptr() + size * int(offsets[s]) for s in seg])
        flat = lambda xc: [xc[i][k] for i, sp in enumerate(trainer.specs) for k in range(len(sp.input_dims))]
        self.s_x_tr = _lib.ptr_table([t.data_ptr() for t in flat(train_xc)])
        self.s_x_te = _lib.ptr_table([t.data_ptr() for t in flat(test_xc)])
        self.s_hat_tr, self.s_hat_te = ptrs(self.xhat_train, self.o_hat_tr, 4), ptrs(self.xhat_test, self.o_hat_te, 4)
        self.s_stats, self.s_roi, self.s_z = ptrs(self.stats, self.o_stats, 4), ptrs(self.roi, self.o_hat_te, 4), ptrs(self.z, self.o_hat_te, 4)
        self.s_subj, self.s_auc_roi = ptrs(self.subj, self.o_subj, 4), ptrs(self.auc_roi, self.o_auc, 8)
        self.s_auc_subj = _lib.ptr_table([self.auc_subj.data_ptr() + 8 * s for s in seg])
        self.s_mask = _lib.ptr_table([None if masks[seg_member[s]] is None else masks[seg_member[s]].data_ptr() for s in seg])
        self.s_lab = _lib.ptr_table([labels[seg_member[s]].data_ptr() for s in seg])
        self.s_ldx, self.s_d = _lib.int_table(seg_ldx), _lib.int_table(seg_d)
        self.s_ntr, self.s_nte, self.s_one = _lib.int_table(n_tr), _lib.int_table(n_te), _lib.int_table([1] * self.n_seg)
        self.launches_per_run = 5

    def run(self):
        lib, st = self.lib, _stream_ptr(self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(lib.nmb_ensemble_reconstruct_sets(self.tr.handle, 2, self.t_xc_both, self.t_rows_both, self.mode, None,
                                                         self.t_hat_both, None, None, st), "nmb_ensemble_reconstruct_sets")
            _lib.check(lib.nmb_normative_stats(self.n_seg, self.s_x_tr, self.s_ldx, self.s_hat_tr, self.s_mask, self.s_ntr,
                                               self.s_d, self.s_stats, st), "nmb_normative_stats")
            _lib.check(lib.nmb_deviation(self.n_seg, self.s_x_te, self.s_ldx, self.s_hat_te, self.s_stats, self.s_nte,
                                         self.s_d, self.s_roi, self.s_z, self.s_subj, st), "nmb_deviation")
            _lib.check(lib.nmb_auc(self.n_seg, self.s_z, self.s_lab, self.s_nte, self.s_d, self.s_auc_roi, None, st), "nmb_auc")
            _lib.check(lib.nmb_auc(self.n_seg, self.s_subj, self.s_lab, self.s_nte, self.s_one, self.s_auc_subj, None, st),
                       "nmb_auc")
        self.tr.gpu_launches += self.launches_per_run
        return self

    def run_deviation_only(self):
        """Just the streaming deviation / z-score kernel on the current reconstructions (for timing it alone)."""
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.nmb_deviation(self.n_seg, self.s_x_te, self.s_ldx, self.s_hat_te, self.s_stats, self.s_nte,
                                              self.s_d, self.s_roi, self.s_z, self.s_subj, _stream_ptr(self.dev)), "nmb_deviation")
        self.tr.gpu_launches += 1
        return self

    # ---- views of segment s ----------------------------------------------------------------
    def seg_stats(self, s):
        return self.stats[int(self.o_stats[s]):int(self.o_stats[s]) + 2 * self.seg_d[s]].view(2, self.seg_d[s])

    def seg_roi(self, s):
        return self.roi[int(self.o_hat_te[s]):int(self.o_hat_te[s]) + self.n_test[s] * self.seg_d[s]].view(self.n_test[s], self.seg_d[s])

    def seg_z(self, s):
        return self.z[int(self.o_hat_te[s]):int(self.o_hat_te[s]) + self.n_test[s] * self.seg_d[s]].view(self.n_test[s], self.seg_d[s])

    def seg_subj(self, s):
        return self.subj[int(self.o_subj[s]):int(self.o_subj[s]) + self.n_test[s]]

    def seg_auc_roi(self, s):
        return self.auc_roi[int(self.o_auc[s]):int(self.o_auc[s]) + self.seg_d[s]]

    def member_table(self, d_max: int = None, n_test_max: int = None) -> torch.Tensor:
        """[n_seg, 1 + 3 d_max + n_test_max] float64: ``member_records`` plus the per-subject deviations (NaN-padded), in
        ONE launch (``nmb_member_records``) -- the rows a rank contributes to the all-gather."""
        import numpy as np
        dmax = max(max(self.seg_d), d_max or 0)
        nmax = max(max(self.n_test), n_test_max or 0)
        if not hasattr(self, "_tbl_dev"):
            t64 = lambda a: torch.from_numpy(np.asarray(a[: self.n_seg], dtype=np.int64)).to(self.dev)
            t32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int32)).to(self.dev)
            self._tbl_dev = (t64(self.o_stats), t64(self.o_auc), t64(self.o_subj), t32(self.seg_d), t32(self.n_test))
        o_stats, o_auc, o_subj, seg_d, n_test = self._tbl_dev
        out = torch.empty((self.n_seg, 1 + 3 * dmax + nmax), dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.nmb_member_records(self.n_seg, self.stats.data_ptr(), o_stats.data_ptr(), self.auc_roi.data_ptr(),
                                                   o_auc.data_ptr(), self.auc_subj.data_ptr(), self.subj.data_ptr(), o_subj.data_ptr(),
                                                   seg_d.data_ptr(), n_test.data_ptr(), dmax, nmax, out.data_ptr(),
                                                   _stream_ptr(self.dev)), "nmb_member_records")
        return out

    def member_records(self, d_max: int = None) -> torch.Tensor:
        """Fixed-size float64 record per segment {subject AUC | per-ROI mean | per-ROI std | per-ROI AUC}, padded
        to the widest segment (or `d_max`, the widest of the WHOLE ensemble when it is sharded over ranks) -- the
        payload of the multi-GPU all-gather (distributed.gather_member_tables)."""
        import numpy as np
        if not hasattr(self, "_rec_idx") or self._rec_dmax != d_max:
            dmax = max(max(self.seg_d), d_max or 0)
            self._rec_dmax = d_max
            w = 1 + 3 * dmax
            src_s, dst_s, src_a, dst_a = [], [], [], []
            for s, d in enumerate(self.seg_d):
                cols = np.arange(d)
                src_s += [self.o_stats[s] + cols, self.o_stats[s] + d + cols]
                dst_s += [s * w + 1 + cols, s * w + 1 + dmax + cols]
                src_a.append(self.o_auc[s] + cols); dst_a.append(s * w + 1 + 2 * dmax + cols)
            t = lambda a: torch.from_numpy(np.concatenate(a).astype(np.int64)).to(self.dev)
            self._rec_idx = (w, t(src_s), t(dst_s), t(src_a), t(dst_a),
                             torch.arange(self.n_seg, device=self.dev, dtype=torch.long) * w)
        w, src_s, dst_s, src_a, dst_a, dst0 = self._rec_idx
        rec = torch.zeros((self.n_seg * w,), dtype=torch.float64, device=self.dev)
        rec.index_copy_(0, dst_s, self.stats.index_select(0, src_s).double())
        rec.index_copy_(0, dst_a, self.auc_roi.index_select(0, src_a))
        rec.index_copy_(0, dst0, self.auc_subj)
        return rec.view(self.n_seg, w)
