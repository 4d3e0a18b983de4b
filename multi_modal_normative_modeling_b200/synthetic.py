"""Synthetic HCP-shaped data: the reference's Google-Drive data is not
available offline, so benchmarks and end-to-end tests use this generator.  Files are written in
the reference layout (``data/<R>/y.csv`` with IID, DIA, AGE, PTGENDER[, FI] and
``data/<R>/<modality>.csv`` with IID + ROI columns) so the reference CLIs could read the same bytes.
"""
from __future__ import annotations

import os
from typing import Dict, Sequence

import numpy as np
import pandas as pd


def make_subjects(n: int = 1000, hc_fraction: float = 0.7, seed: int = 42) -> pd.DataFrame:
    """IID sub-%04d; DIA 1 = healthy control (HCPimage convention, utils.py:770-771), 0 = patient."""
    rs = np.random.RandomState(seed)
    n_hc = int(round(n * hc_fraction))
    dia = np.zeros(n, dtype=np.int64)
    dia[rs.permutation(n)[:n_hc]] = 1
    return pd.DataFrame({
        "IID": [f"sub-{i:04d}" for i in range(n)],
        "DIA": dia,
        "AGE": rs.randint(22, 37, n).astype(np.int64),
        "PTGENDER": rs.randint(1, 3, n).astype(np.int64),
        "FI": rs.normal(105.0, 15.0, n),
    })


def make_modality(subjects: pd.DataFrame, d: int, seed: int, n_factors: int = 8, effect: float = 0.8,
                  affected_fraction: float = 0.2) -> np.ndarray:
    """X = a*age_z + b*sex + L f + noise; patients get +effect*sigma on a fixed 20 % of ROIs;
    then a per-modality affine so that RobustScaler matters."""
    rs = np.random.RandomState(seed)
    n = len(subjects)
    age = subjects["AGE"].to_numpy(np.float64)
    age_z = (age - age.mean()) / (age.std() + 1e-12)
    sex = subjects["PTGENDER"].to_numpy(np.float64) - 1.5
    a, b = rs.normal(0, 0.3, d), rs.normal(0, 0.3, d)
    load = rs.normal(0, 0.5, (n_factors, d))
    f = rs.normal(0, 1, (n, n_factors))
    x = np.outer(age_z, a) + np.outer(sex, b) + f @ load + rs.normal(0, 1, (n, d))
    affected = rs.permutation(d)[: max(1, int(d * affected_fraction))]
    patients = subjects["DIA"].to_numpy() == 0
    x[np.ix_(patients, affected)] += effect * x.std(axis=0)[affected]
    scale = 10.0 ** rs.uniform(1, 3)
    return x * scale + rs.normal(0, scale, d)


def make_hcpimage(n: int = 1000, d: int = 116, modalities: Sequence[str] = ("T1w_sMRI", "T2w_sMRI", "fMRI"),
                  seed: int = 42) -> Dict[str, object]:
    subjects = make_subjects(n, seed=seed)
    feats = {m: make_modality(subjects, d, seed + 1 + i) for i, m in enumerate(modalities)}
    return {"subjects": subjects, "features": feats}


def write_dataset(root: str, resource: str = "HCPimage", n: int = 1000, seed: int = 42, columns=None) -> str:
    """Write data/<resource>/{y,<modality>,early_fusion_modalities_<resource>}.csv under root."""
    from .utils import COLUMNS_NAME_AAL116
    cols = list(columns or COLUMNS_NAME_AAL116)
    data = make_hcpimage(n, len(cols), seed=seed)
    out = os.path.join(root, "data", resource)
    os.makedirs(out, exist_ok=True)
    data["subjects"].to_csv(os.path.join(out, "y.csv"), index=False)
    parts = [pd.DataFrame({"IID": data["subjects"]["IID"]})]
    for name, x in data["features"].items():
        df = pd.DataFrame(x, columns=cols)
        df.insert(0, "IID", data["subjects"]["IID"].to_numpy())
        df.to_csv(os.path.join(out, name + ".csv"), index=False)
        parts.append(pd.DataFrame(x, columns=[f"{c}_{name}" for c in cols]))   # early_fusion_modalities.py:10-35
    fused = pd.concat(parts, axis=1)
    fused.to_csv(os.path.join(out, f"early_fusion_modalities_{resource}.csv"), index=False)
    return out
