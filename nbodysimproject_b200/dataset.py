"""Feature table <-> the on-disk dataset format the reference's trainers consume.

Reader: `StabilityDataset.load / get_metadata` keep the reference's semantics (stability_dataset.py:21-125): an optional
first line `# feature_names: a,b,c`, a pandas-readable CSV body, `is_stable` as the label, `simulation_id`, `mode`,
`dataset_version` and every `scaler_mean_* / scaler_scale_*` column excluded from the feature matrix, rows with a NaN
label dropped, NaN features replaced by 0.

Writer: the reference has no writer for the header / scaler columns (only `DataFrame.to_csv`,
batch_stability_analyzer.py:82-88); `save_feature_table` writes the full format straight from the feature tensors that
`nb_ensemble_run_f64` / `nb_ensemble_prepare_f64` produce (dyn_features[B][22] + static_features[B][25]) without
building one Python dict per system, which at B = 10^6 is the dominant cost of the reference's path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L

NON_FEATURE_COLUMNS = ["simulation_id", "is_stable", "mode", "dataset_version"]
# the 17 dynamic columns a user sees (stability_analyzer.py:226-252); the _E0.._t_end taps are internal
PUBLIC_DYN = L.DYN_COLUMNS[:17]


def feature_columns(analysis_mode: str = "full") -> List[str]:
    """Column order of BatchStabilityAnalyzer.analyze_batch(...) for `analysis_mode`."""
    cols = list(PUBLIC_DYN) + ["mode"]
    if analysis_mode == "full":
        cols += ["initial_" + c for c in L.STATIC_COLUMNS]
    return cols + ["pathological_energy", "softening_policy", "simulation_id"]


def table_from_tensors(dyn, static=None, analysis_mode: str = "full", softening_policy: str = "static",
                       first_id: int = 0):
    """pandas DataFrame with the reference's columns from the raw feature tensors (torch or numpy), vectorised."""
    import pandas as pd
    dyn = dyn.detach().cpu().numpy() if hasattr(dyn, "detach") else np.asarray(dyn)
    data: Dict[str, np.ndarray] = {c: dyn[:, i] for i, c in enumerate(PUBLIC_DYN)}
    B = dyn.shape[0]
    data["mode"] = np.full(B, analysis_mode, dtype=object)
    if analysis_mode == "full" and static is not None:
        st = static.detach().cpu().numpy() if hasattr(static, "detach") else np.asarray(static)
        for i, c in enumerate(L.STATIC_COLUMNS):
            data["initial_" + c] = st[:, i]
    e = dyn[:, L.DYN_COLUMNS.index("energy_drift")]
    with np.errstate(invalid="ignore"):
        patho = np.abs(e) > 10.0                                  # batch_stability_analyzer.py:45-52
    data["is_stable"] = np.where(patho, 0.0, data["is_stable"])
    data["pathological_energy"] = patho
    data["softening_policy"] = np.full(B, softening_policy, dtype=object)
    data["simulation_id"] = np.arange(first_id, first_id + B)
    return pd.DataFrame(data)


def numeric_feature_columns(df) -> List[str]:
    """The columns StabilityDataset.load will return as features once `df` is written by save_feature_table."""
    import pandas as pd
    return [c for c in df.columns if c not in NON_FEATURE_COLUMNS
            and (pd.api.types.is_numeric_dtype(df[c]) or pd.api.types.is_bool_dtype(df[c]))]


def save_feature_table(path: str, df, feature_names: Optional[Sequence[str]] = None,
                       scaler_mean: Optional[Sequence[float]] = None, scaler_scale: Optional[Sequence[float]] = None,
                       dataset_version: Optional[str] = None) -> List[str]:
    """Write `df` in the format StabilityDataset.load reads: `# feature_names:` header line, optional
    scaler_mean_<k> / scaler_scale_<k> columns (constant per row; the loader takes row 0), optional dataset_version.
    Non-numeric columns other than the excluded ones are dropped from the feature list (the loader's
    `df[feature_cols].values` must be numeric).  Returns the feature names written to the header."""
    import pandas as pd
    df = df.copy()
    numeric = lambda c: pd.api.types.is_numeric_dtype(df[c]) or pd.api.types.is_bool_dtype(df[c])
    df = df.drop(columns=[c for c in df.columns if c not in NON_FEATURE_COLUMNS and not numeric(c)])
    for c in df.columns:
        if pd.api.types.is_bool_dtype(df[c]):
            df[c] = df[c].astype(float)
    feats = [c for c in df.columns if c not in NON_FEATURE_COLUMNS]
    if feature_names is None:
        feature_names = feats
    if dataset_version is not None:
        df["dataset_version"] = dataset_version
    if scaler_mean is not None and scaler_scale is not None:
        if len(scaler_mean) != len(feats) or len(scaler_scale) != len(feats):
            raise L.NBodyB200Error("scaler_mean / scaler_scale must have one entry per feature column")
        width = len(str(len(feats) - 1))
        for k, (mu, sc) in enumerate(zip(scaler_mean, scaler_scale)):      # zero-padded so sorted() keeps the order
            df[f"scaler_mean_{k:0{width}d}"] = float(mu)
        for k, (mu, sc) in enumerate(zip(scaler_mean, scaler_scale)):
            df[f"scaler_scale_{k:0{width}d}"] = float(sc)
    with open(path, "w") as f:
        f.write("# feature_names: " + ",".join(feature_names) + "\n")
        df.to_csv(f, index=False)
    return list(feature_names)


class StabilityDataset:
    """stability_dataset.py:18-125."""

    @staticmethod
    def load(path: str) -> Tuple[np.ndarray, np.ndarray, List[str]]:
        import pandas as pd
        feature_names = None
        with open(path, "r") as f:
            first = f.readline()
            if first.startswith("# feature_names:"):
                feature_names = first.strip().split(":", 1)[1].strip().split(",")
        df = pd.read_csv(path, comment="#")
        if "is_stable" not in df.columns:
            print("[error] CSV must contain 'is_stable' column")
            return np.array([]), np.array([]), []
        exclude = list(NON_FEATURE_COLUMNS) + [c for c in df.columns if c.startswith("scaler_")]
        cols = [c for c in df.columns if c not in exclude]
        if feature_names is None:
            feature_names = cols
        X = df[cols].values
        y = df["is_stable"].values
        ok = ~np.isnan(y)
        X, y = X[ok], y[ok]
        print(f"Loaded {len(X)} samples with {X.shape[1]} features")
        if np.any(np.isnan(X)):
            print("[warning] NaN values found in features. Replacing with 0.")
            X = np.nan_to_num(X, nan=0.0)
        return X, y, feature_names

    @staticmethod
    def get_metadata(path: str) -> Dict:
        import pandas as pd
        meta = {"feature_names": None, "scaler_mean": None, "scaler_scale": None}
        with open(path, "r") as f:
            first = f.readline()
            if first.startswith("# feature_names:"):
                meta["feature_names"] = first.strip().split(":", 1)[1].strip().split(",")
        df = pd.read_csv(path, comment="#", nrows=1)
        mean_cols = [c for c in df.columns if c.startswith("scaler_mean_")]
        scale_cols = [c for c in df.columns if c.startswith("scaler_scale_")]
        if mean_cols:
            meta["scaler_mean"] = df[mean_cols].iloc[0].values
        if scale_cols:
            meta["scaler_scale"] = df[scale_cols].iloc[0].values
        return meta
