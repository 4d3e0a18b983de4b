"""ctypes binding of libnmb.so (include/nmb.h).  There is no CPU or eager fallback: if the
library is missing or an entry point fails, a ``RuntimeError`` is raised."""
from __future__ import annotations

import ctypes as C
import os

NMB_MAX_MOD = 16
NMB_MAX_HIDDEN = 4
NMB_MAX_HEAD = 3

COMBINE = {"poe": 0, "gpoe": 1, "moe": 2, "mopoe": 3}
LOSS = {"gauss_ll": 0, "neg_mse": 1}
HEAD = {None: 0, "none": 0, "regression": 1, "endtoend": 2}
(SLOT_ENC, SLOT_ENC_MEAN, SLOT_ENC_LOGVAR, SLOT_DEC, SLOT_DEC_MEAN, SLOT_LOGVAR_OUT, SLOT_ALPHA, SLOT_HEAD, SLOT_HEAD_BN,
 SLOT_DEC2, SLOT_DEC2_MEAN, SLOT_LOGVAR_OUT2) = range(12)
TRAIN_NO_ADAM, TRAIN_WRITE_GRADS, TRAIN_KEEP_ACTS, TRAIN_FP32, TRAIN_TC_SIMPLE, TRAIN_RESIDENT, TRAIN_LOSS4 = 1, 2, 4, 8, 16, 32, 64
TRAIN_LOSS8, TRAIN_NO_STATS = 128, 256
HP_MARGIN, HP_W_CONTRASTIVE, HP_W_KL, HP_W_REC, HP_DROPOUT = range(5)
RECON_MEAN, RECON_SAMPLE, RECON_GIVEN_Z, RECON_FP32, RECON_TC_SIMPLE, RECON_KEEP_PLANES = 0, 1, 2, 16, 32, 64

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnmb.so")


class NmbArch(C.Structure):
    _fields_ = [("n_mod", C.c_int32), ("input_dims", C.c_int32 * NMB_MAX_MOD), ("n_hidden", C.c_int32),
                ("hidden", C.c_int32 * NMB_MAX_HIDDEN), ("latent", C.c_int32), ("c_dim", C.c_int32),
                ("combine", C.c_int32), ("loss_kind", C.c_int32), ("non_linear", C.c_int32),
                ("head_kind", C.c_int32), ("n_head_hidden", C.c_int32), ("head_hidden", C.c_int32 * NMB_MAX_HEAD),
                ("head_weight", C.c_float), ("head_params", C.c_float * 6),
                ("family", C.c_int32), ("s_dim", C.c_int32), ("weighted", C.c_int32), ("beta", C.c_float)]


class NmbSlot(C.Structure):
    _fields_ = [("kind", C.c_int32), ("modality", C.c_int32), ("layer", C.c_int32), ("rows", C.c_int32),
                ("cols", C.c_int32), ("ld", C.c_int32), ("offset", C.c_int64)]


class NmbMember(C.Structure):
    _fields_ = [("arch", NmbArch), ("xc", C.c_void_p * NMB_MAX_MOD), ("n_rows", C.c_int32), ("batch", C.c_int32),
                ("seed", C.c_uint64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("adam_eps", C.c_float), ("lr_steps", C.c_void_p), ("params", C.c_void_p),
                ("adam_m", C.c_void_p), ("adam_v", C.c_void_p), ("grads", C.c_void_p), ("n_lr_steps", C.c_int64),
                ("y", C.c_void_p), ("row_order", C.c_void_p), ("n_order_epochs", C.c_int64),
                ("drop_keep", C.c_void_p), ("n_drop_steps", C.c_int64)]


_PROTOS = {
    "nmb_last_error": (C.c_char_p, []),
    "nmb_version": (C.c_int, []),
    "nmb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "nmb_arch_param_count": (C.c_int, [C.POINTER(NmbArch), C.POINTER(C.c_int64)]),
    "nmb_arch_slots": (C.c_int, [C.POINTER(NmbArch), C.POINTER(NmbSlot), C.c_int32, C.POINTER(C.c_int32)]),
    "nmb_packed_row_stride": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "nmb_pack_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "nmb_robust_fit": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nmb_rank_bins": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "nmb_pack_rows_scaled": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "nmb_csv_write": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                C.c_int64, C.c_int32]),
    "nmb_ensemble_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(NmbMember), C.c_int32]),
    "nmb_ensemble_destroy": (C.c_int, [C.c_void_p]),
    "nmb_ensemble_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "nmb_ensemble_engine": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32)]),
    "nmb_ensemble_steps_done": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "nmb_ensemble_train": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "nmb_ensemble_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nmb_ensemble_invalidate": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nmb_ensemble_train_epochs": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_uint32, C.c_void_p]),
    "nmb_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "nmb_ensemble_peek": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int32), C.c_void_p]),
    "nmb_ensemble_peek_head": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "nmb_ensemble_reconstruct": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32,
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.c_void_p]),
    "nmb_ensemble_head_predict": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32,
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "nmb_ensemble_reconstruct_sets": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32,
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                C.POINTER(C.c_void_p), C.c_void_p]),
    "nmb_normative_stats": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_void_p), C.c_void_p]),
    "nmb_deviation": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "nmb_latent_deviation": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "nmb_auc": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                          C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "nmb_member_records": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "nmb_mean_rows": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "nmb_philox_normal": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]),
    "nmb_debug_tcp_trace": (C.c_int, [C.c_void_p, C.c_int32]),
    "nmb_debug_tc_gemm": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
}

EXPORTS = tuple(_PROTOS)
_lib = None


def load():
    """Load libnmb.so (built in-tree by ``_build.build()``).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m multi_modal_normative_modeling_b200._build` "
            "(the CUDA extension is mandatory; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status: int, what: str = "libnmb"):
    if status != 0:
        msg = load().nmb_last_error().decode("utf-8", "replace")
        if "No such combination method" in msg:
            raise ValueError("No such combination method")      # cVAE.py:1163
        raise RuntimeError(f"{what} failed ({status}): {msg}")


def ptr_table(ptrs):
    """Host array of device pointers (None -> NULL)."""
    arr = (C.c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


def int_table(vals):
    return (C.c_int32 * max(len(vals), 1))(*[int(v) for v in vals])


def make_arch(input_dims, hidden, latent, c_dim, combine="poe", loss_kind="gauss_ll", non_linear=True,
              head=None, head_hidden=(128, 64), head_weight=1.0, head_params=None,
              family=None, s_dim=0, weighted=False, beta=1.0) -> NmbArch:
    if isinstance(combine, str):
        key = combine.lower()
        if key not in COMBINE:
            raise ValueError("No such combination method")     # cVAE.py:1163
        combine = COMBINE[key]
    if isinstance(loss_kind, str):
        loss_kind = LOSS[loss_kind]
    if len(input_dims) > NMB_MAX_MOD:
        raise ValueError(f"at most {NMB_MAX_MOD} modalities")
    if not 1 <= len(hidden) <= NMB_MAX_HIDDEN:
        raise ValueError(f"1..{NMB_MAX_HIDDEN} hidden layers supported")
    a = NmbArch()
    a.n_mod = len(input_dims)
    for i, d in enumerate(input_dims):
        a.input_dims[i] = int(d)
    a.n_hidden = len(hidden)
    for i, h in enumerate(hidden):
        a.hidden[i] = int(h)
    a.latent, a.c_dim = int(latent), int(c_dim)
    a.combine, a.loss_kind, a.non_linear = int(combine), int(loss_kind), int(bool(non_linear))
    a.head_kind = HEAD[head.lower() if isinstance(head, str) else head]
    if a.head_kind:
        if not 1 <= len(head_hidden) <= NMB_MAX_HEAD:
            raise ValueError(f"1..{NMB_MAX_HEAD} hidden layers in a supervised head")
        a.n_head_hidden = len(head_hidden)
        for i, h in enumerate(head_hidden):
            a.head_hidden[i] = int(h)
        a.head_weight = float(head_weight)
        if a.head_kind == HEAD["endtoend"]:
            hp = dict(margin=1.0, w_contrastive=0.1, w_kl=0.1, w_rec=0.1, dropout=0.5)     # cVAE.py:2131, 2031
            hp.update(head_params or {})
            for i, k in enumerate(("margin", "w_contrastive", "w_kl", "w_rec", "dropout")):
                a.head_params[i] = float(hp[k])
            a.head_weight = 0.0
    if family in ("dmvae", 1):         # DMVAE / mmVAEPlus / WeightedDMVAE (cVAE.py:1491-1752, 1895-2002)
        a.family, a.s_dim, a.weighted, a.beta = 1, int(s_dim), int(bool(weighted)), float(beta)
    elif family in ("mvtcae", 2):      # mvtCAE (cVAE.py:1754-1893): beta weighs the total-correlation term
        a.family, a.beta = 2, float(beta)
    elif family not in (None, "cvae", 0):
        raise ValueError("unknown model family")
    return a


def arch_slots(arch: NmbArch):
    lib = load()
    n = C.c_int32(0)
    check(lib.nmb_arch_slots(C.byref(arch), None, 0, C.byref(n)), "nmb_arch_slots")
    slots = (NmbSlot * n.value)()
    check(lib.nmb_arch_slots(C.byref(arch), slots, n.value, C.byref(n)), "nmb_arch_slots")
    return list(slots)


def arch_param_count(arch: NmbArch) -> int:
    n = C.c_int64(0)
    check(load().nmb_arch_param_count(C.byref(arch), C.byref(n)), "nmb_arch_param_count")
    return n.value


def packed_row_stride(d: int, c_dim: int) -> int:
    n = C.c_int32(0)
    check(load().nmb_packed_row_stride(d, c_dim, C.byref(n)), "nmb_packed_row_stride")
    return n.value
