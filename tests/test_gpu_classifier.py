"""GPU numerics: the fused scaler + MLP + sigmoid + threshold kernel against a plain PyTorch fp32 reference of the same
op built from the reference's architecture (model_zoo.py:18-33).  Tolerance 2e-5 absolute on the probability (fp32,
different summation order); labels must agree wherever the probability is not within 1e-4 of the threshold."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,B", [("full", 5000), ("core", 77), ("full", 1)])
def test_mlp_matches_torch_fp32(mode, B):
    import torch
    from nbodysimproject_b200 import _lib as L
    from nbodysimproject_b200.classifier import StabilityClassifier, default_feature_index
    torch.manual_seed(0)
    names, idx = default_feature_index(mode)
    F = len(names)
    model = torch.nn.Sequential()       # same layers / names as model_zoo.MLP
    fc1, fc2, fc3 = torch.nn.Linear(F, 128), torch.nn.Linear(128, 64), torch.nn.Linear(64, 1)
    sd = {"fc1.weight": fc1.weight, "fc1.bias": fc1.bias, "fc2.weight": fc2.weight, "fc2.bias": fc2.bias,
          "fc3.weight": fc3.weight, "fc3.bias": fc3.bias}
    rng = np.random.default_rng(1)
    dyn = rng.standard_normal((B, L.N_DYN)) * 3.0
    stat = rng.standard_normal((B, L.N_STATIC)) * 3.0
    dyn[::7, 5] = np.nan                                    # nan_to_num path
    dyn[::11, 1] = 40.0                                     # pathological energy drift -> derived column = 1
    mean, scale = rng.standard_normal(F), rng.uniform(0.5, 2.0, F)
    clf = StabilityClassifier.from_state_dict(sd, mean=mean, scale=scale, threshold=0.37, feature_index=idx)
    d_dyn, d_stat = torch.as_tensor(dyn).cuda(), torch.as_tensor(stat).cuda()
    prob, label = clf.predict(d_dyn, d_stat if mode == "full" else None)
    # torch fp32 reference
    patho = (np.abs(dyn[:, L.DYN_COLUMNS.index("energy_drift")]) > 10.0).astype(float)
    wide = np.concatenate([dyn, stat, patho[:, None]], axis=1)
    X = wide[:, [c if c < 63 else (L.N_DYN + L.N_STATIC if c == 63 else L.N_DYN + (c - 64)) for c in idx]]
    X = np.nan_to_num(X, nan=0.0)
    Xs = torch.as_tensor(((X.astype(np.float32) - mean.astype(np.float32)) * (1.0 / scale).astype(np.float32)))
    with torch.no_grad():
        ref = torch.sigmoid(fc3(torch.relu(fc2(torch.relu(fc1(Xs)))))).reshape(-1).numpy()
    got = prob.cpu().numpy()
    assert np.max(np.abs(got - ref)) < 2e-5
    sure = np.abs(ref - 0.37) > 1e-4
    assert np.array_equal(label.cpu().numpy()[sure], (ref > 0.37).astype(np.int32)[sure])


def test_classifier_on_real_feature_tensors():
    """End of the loop: ensemble analysis -> feature tensors -> classifier, all on the device."""
    import torch
    from nbodysimproject_b200 import _lib as L, ensemble as E
    from nbodysimproject_b200.classifier import StabilityClassifier, default_feature_index
    from nbodysimproject_b200.generators import EnsembleInputs
    rng = np.random.default_rng(5)
    m, q, v, soft, _ = EnsembleInputs.diverse(rng, 4096, n_max=5)[4]
    B = m.shape[0]
    bk = E.DeviceBucket(m, q, v, soft, 1.0, "yoshida4")
    bk.prepare(L.PREP_REMOVE_COM | L.PREP_CTOR_KICK | L.PREP_SNAPSHOT_KICK, 0.01, 0.01, 0.01, 50, want_static=True)
    bk.sort()
    dyn = bk.run(0.01, 100, 1, 50, rng.standard_normal((B, 4, 2)), rng.standard_normal((B, 4, 2)), flags=L.RUN_ENERGY)
    names, idx = default_feature_index("full")
    torch.manual_seed(1)
    F = len(names)
    sd = {"fc1.weight": torch.randn(128, F) * 0.1, "fc1.bias": torch.zeros(128), "fc2.weight": torch.randn(64, 128) * 0.1,
          "fc2.bias": torch.zeros(64), "fc3.weight": torch.randn(1, 64) * 0.1, "fc3.bias": torch.zeros(1)}
    clf = StabilityClassifier.from_state_dict(sd, feature_index=idx)
    prob, label = clf.predict(dyn, bk.static)
    p = prob.cpu().numpy()
    assert p.shape == (B,) and np.all(np.isfinite(p[np.isfinite(dyn.cpu().numpy()).all(1)]))
    assert set(np.unique(label.cpu().numpy())) <= {0, 1}


def test_reference_trained_layout_round_trip(tmp_path):
    """A model trained by the reference's train_mlp on a table written from the feature tensors has one input per column
    StabilityDataset.load returns (42 in full mode: 16 dynamic + 25 static + pathological_energy).  Its state_dict must be
    accepted with the DEFAULT feature index, and the kernel's gather must line up with the loader's matrix and the
    scaler vectors column for column (the derived column was missing from the index)."""
    import torch
    from nbodysimproject_b200 import _lib as L, dataset as D
    from nbodysimproject_b200.classifier import StabilityClassifier, default_feature_index
    rng = np.random.default_rng(3)
    B = 600
    dyn = rng.standard_normal((B, L.N_DYN)) * 2.0
    dyn[:, 0] = (rng.random(B) > 0.5).astype(float)
    dyn[::9, 1] = 30.0
    dyn[::13, 6] = np.nan
    stat = rng.standard_normal((B, L.N_STATIC)) * 2.0
    df = D.table_from_tensors(dyn, stat, "full")
    path = str(tmp_path / "t.csv")
    names = D.save_feature_table(path, df)
    X, y, fn = D.StabilityDataset.load(path)
    assert X.shape == (B, 42) and fn == names == default_feature_index("full")[0]
    mean, scale = X.mean(0), X.std(0) + 0.5
    torch.manual_seed(4)
    fc1, fc2, fc3 = torch.nn.Linear(42, 128), torch.nn.Linear(128, 64), torch.nn.Linear(64, 1)
    sd = {"fc1.weight": fc1.weight, "fc1.bias": fc1.bias, "fc2.weight": fc2.weight, "fc2.bias": fc2.bias,
          "fc3.weight": fc3.weight, "fc3.bias": fc3.bias}
    clf = StabilityClassifier.from_state_dict(sd, mean=mean, scale=scale)          # default index
    prob, _ = clf.predict(torch.as_tensor(dyn).cuda(), torch.as_tensor(stat).cuda())
    Xs = torch.as_tensor(((X.astype(np.float32) - mean.astype(np.float32)) * (1.0 / scale).astype(np.float32)))
    with torch.no_grad():
        ref = torch.sigmoid(fc3(torch.relu(fc2(torch.relu(fc1(Xs)))))).reshape(-1).numpy()
    assert np.max(np.abs(prob.cpu().numpy() - ref)) < 5e-5
