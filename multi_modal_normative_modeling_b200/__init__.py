"""B200-native (sm_100a) implementation of the cVAE-ensemble hot path of
soz223/multi_modal_normative_modeling: batched training of many small conditional VAEs and
per-subject / per-ROI deviation scoring, behind the reference's own Python API."""
from . import _lib  # noqa: F401
from .ensemble import EnsembleTrainer, MemberSpec, pack_rows  # noqa: F401

__all__ = ["EnsembleTrainer", "MemberSpec", "pack_rows"]
