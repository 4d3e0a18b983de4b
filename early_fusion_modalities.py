#!/usr/bin/env python3
"""Column-wise concatenation of the modality CSVs into early_fusion_modalities_<R>.csv with
``_<modality>`` column suffixes (reference early_fusion_modalities.py:10-35)."""
import argparse
from pathlib import Path

import pandas as pd

from multi_modal_normative_modeling_b200.utils import get_column_name, get_datasets_name


def main(resource, root="."):
    base = Path(root) / "data" / resource
    parts = []
    for i, name in enumerate(get_datasets_name(resource)):
        df = pd.read_csv(base / f"{name}.csv")
        cols = get_column_name(resource, name)
        body = df[cols].rename(columns={c: f"{c}_{name}" for c in cols})
        parts.append(pd.concat([df[["IID"]], body], axis=1) if i == 0 else body)
    pd.concat(parts, axis=1).to_csv(base / f"early_fusion_modalities_{resource}.csv", index=False)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-R", "--dataset_resourse", default="ADNI")
    main(ap.parse_args().dataset_resourse)
