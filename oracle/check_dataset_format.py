"""Build-container check (needs /root/reference): a file written by nbodysimproject_b200.dataset.save_feature_table is
read by the REFERENCE's StabilityDataset.load / get_metadata with the same result as by our loader.  Also (re)writes
tests/golden/dataset_sample.csv."""
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
sys.modules.setdefault("lightgbm", types.ModuleType("lightgbm"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import minbody as mb  # noqa: E402
from nbodysimproject_b200 import dataset as D  # noqa: E402
from test_dataset_format import _tensors  # noqa: E402

dyn, static = _tensors()
df = D.table_from_tensors(dyn, static, "full")
n_feat = len(D.numeric_feature_columns(df))
path = os.path.join(ROOT, "tests", "golden", "dataset_sample.csv")
names = D.save_feature_table(path, df, scaler_mean=np.arange(n_feat), scaler_scale=np.arange(n_feat) + 1.0,
                             dataset_version="r1")
Xr, yr, fr = mb.StabilityDataset.load(path)
Xo, yo, fo = D.StabilityDataset.load(path)
assert fr == fo == names, (fr, fo)
assert np.array_equal(Xr, Xo) and np.array_equal(yr, yo)
mr, mo = mb.StabilityDataset.get_metadata(path), D.StabilityDataset.get_metadata(path)
assert mr["feature_names"] == mo["feature_names"]
assert np.array_equal(mr["scaler_mean"], mo["scaler_mean"]) and np.array_equal(mr["scaler_scale"], mo["scaler_scale"])
assert np.array_equal(mr["scaler_mean"], np.arange(n_feat))
print("reference loader and ours agree:", Xr.shape, len(fr), "features; scaler columns in order")
