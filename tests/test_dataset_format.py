"""CPU: feature tensors -> on-disk dataset format -> loader round trip (SURVEY.md section 8f item 2;
stability_dataset.py:21-125).  tests/golden/dataset_sample.csv was written by save_feature_table and verified once
against the REFERENCE's own StabilityDataset.load / get_metadata in the build container (oracle/check_dataset_format.py)."""
import os

import numpy as np

from conftest import GOLDEN


def _tensors(B=7, seed=0):
    from nbodysimproject_b200 import _lib as L
    rng = np.random.default_rng(seed)
    dyn = rng.standard_normal((B, L.N_DYN))
    dyn[:, 0] = (rng.random(B) > 0.5).astype(float)
    dyn[2, 1] = 50.0                 # pathological energy drift -> is_stable forced to 0
    dyn[3, 0] = np.nan               # unlabeled row: dropped by the loader
    dyn[4, 5] = np.nan               # NaN feature: replaced by 0
    static = rng.standard_normal((B, L.N_STATIC))
    return dyn, static


def test_table_columns_match_reference_order():
    from nbodysimproject_b200 import dataset as D
    dyn, static = _tensors()
    df = D.table_from_tensors(dyn, static, "full")
    assert list(df.columns) == D.feature_columns("full")
    assert len(df.columns) == 46
    assert df["is_stable"][2] == 0.0 and bool(df["pathological_energy"][2])
    core = D.table_from_tensors(dyn, None, "core")
    assert list(core.columns) == D.feature_columns("core")


def test_round_trip_through_the_loader(tmp_path):
    from nbodysimproject_b200 import dataset as D
    dyn, static = _tensors()
    df = D.table_from_tensors(dyn, static, "full")
    n_feat = len(D.numeric_feature_columns(df))
    mean, scale = np.arange(n_feat, dtype=float), np.arange(n_feat, dtype=float) + 1.0
    path = str(tmp_path / "d.csv")
    names = D.save_feature_table(path, df, scaler_mean=mean, scaler_scale=scale, dataset_version="r1")
    assert open(path).readline().startswith("# feature_names: energy_drift,")
    X, y, fn = D.StabilityDataset.load(path)
    assert fn == names and X.shape == (6, n_feat) and y.shape == (6,)       # the NaN-label row is gone
    assert not np.isnan(X).any()
    keep = [i for i in range(7) if i != 3]
    assert np.allclose(X[:, 0], dyn[keep, 1])                              # first feature = energy_drift
    meta = D.StabilityDataset.get_metadata(path)
    assert meta["feature_names"] == names
    assert np.array_equal(meta["scaler_mean"], mean) and np.array_equal(meta["scaler_scale"], scale)


def test_committed_sample_loads():
    from nbodysimproject_b200 import dataset as D
    X, y, fn = D.StabilityDataset.load(os.path.join(GOLDEN, "dataset_sample.csv"))
    assert X.shape[0] == y.shape[0] == 6 and X.shape[1] == len(fn)


def test_classifier_feature_index_matches_the_written_table(tmp_path):
    """The gather map of the inference kernel lists exactly the columns the loader returns, in the loader's order."""
    from nbodysimproject_b200 import dataset as D
    from nbodysimproject_b200.classifier import default_feature_index, COL_PATHOLOGICAL
    dyn, static = _tensors()
    for mode, n in (("full", 42), ("core", 17)):
        df = D.table_from_tensors(dyn, static if mode == "full" else None, mode)
        path = str(tmp_path / f"{mode}.csv")
        names = D.save_feature_table(path, df)
        X, y, fn = D.StabilityDataset.load(path)
        idx_names, idx = default_feature_index(mode)
        assert fn == names == idx_names and X.shape[1] == n == len(idx)
        assert idx[-1] == COL_PATHOLOGICAL and names[-1] == "pathological_energy"
