"""Stability-classifier inference fused onto the GPU feature tensors.

`StabilityClassifier` holds the weights of the reference's MLP (model_zoo.py:18-33), its StandardScaler
(train_mlp.py:52-60, scaler_utils.py:18-27) and decision threshold (train_mlp.py:141-187) on the device and classifies
the `dyn_features` / `static_features` tensors of an ensemble analysis in one kernel (`nb_mlp_classify_f32`).
Training stays out of scope; weights come from a torch state_dict saved by the reference's trainer."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from .dataset import PUBLIC_DYN


COL_PATHOLOGICAL = 63      # NB_MLP_COL_PATHOLOGICAL: |energy_drift| > 10, synthesised by the kernel


def default_feature_index(analysis_mode: str = "full"):
    """Feature order of StabilityDataset.load on a table written by dataset.save_feature_table (which is the order the
    reference's loader, stability_dataset.py:21-125, sees in a DataFrame.to_csv of analyze_batch): the public dynamic
    columns minus the label, then (full mode) the 25 static columns, then `pathological_energy` -- a numeric column of
    the table, so a model trained by the reference's train_mlp on such a file has 42 (full) / 17 (core) inputs.
    Returns (names, index): index < 63 -> dyn column, 63 -> pathological_energy, 64 + c -> static column c."""
    names, idx = [], []
    for c, name in enumerate(PUBLIC_DYN):
        if name == "is_stable":
            continue
        names.append(name); idx.append(c)
    if analysis_mode == "full":
        for c, name in enumerate(L.STATIC_COLUMNS):
            names.append("initial_" + name); idx.append(64 + c)
    names.append("pathological_energy"); idx.append(COL_PATHOLOGICAL)
    return names, np.asarray(idx, dtype=np.int32)


class StabilityClassifier:
    def __init__(self, w1, b1, w2, b2, w3, b3, mean=None, scale=None, threshold: float = 0.5,
                 feature_index: Optional[Sequence[int]] = None, device=None):
        """w1 [128, F], w2 [64, 128], w3 [1, 64] or [64] in torch's nn.Linear layout (out_features first)."""
        torch = L.require_cuda()
        self.torch = torch
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        w1 = np.asarray(w1, dtype=np.float32); w2 = np.asarray(w2, dtype=np.float32)
        w3 = np.asarray(w3, dtype=np.float32).reshape(-1)
        if w1.shape[0] != 128 or w2.shape != (64, 128) or w3.shape != (64,):
            raise L.NBodyB200Error("expected the reference's MLP: F -> 128 -> 64 -> 1 (model_zoo.py:18-33)")
        self.F = int(w1.shape[1])
        if feature_index is None:
            _, feature_index = default_feature_index("full" if self.F > 17 else "core")
            if feature_index.size == self.F + 1:              # a table written without the derived column
                feature_index = feature_index[:-1]
        feature_index = np.asarray(feature_index, dtype=np.int32)
        if feature_index.size != self.F or self.F > 64:
            raise L.NBodyB200Error(f"feature_index must list the {self.F} input columns (F <= 64)")
        mean = np.zeros(self.F) if mean is None else np.asarray(mean, dtype=np.float64)
        scale = np.ones(self.F) if scale is None else np.asarray(scale, dtype=np.float64)
        dev = self.device
        t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a)).to(dev, dt).contiguous()
        self.idx = t(feature_index, torch.int32)
        self.mean, self.inv_scale = t(mean), t(1.0 / scale)
        self.w1, self.b1 = t(w1.T), t(np.asarray(b1, dtype=np.float32))       # input-major for the kernel
        self.w2, self.b2 = t(w2.T), t(np.asarray(b2, dtype=np.float32))
        self.w3, self.b3 = t(w3), float(np.asarray(b3).reshape(-1)[0])
        self.threshold = float(threshold)

    @classmethod
    def from_state_dict(cls, sd, **kw):
        g = lambda k: sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else np.asarray(sd[k])
        return cls(g("fc1.weight"), g("fc1.bias"), g("fc2.weight"), g("fc2.bias"), g("fc3.weight"), g("fc3.bias"), **kw)

    def predict(self, dyn, static=None):
        """dyn [B, 22] (and static [B, 25]) fp64 device tensors -> (prob [B] fp32, label [B] int32), on the device."""
        torch = self.torch
        if bool((self.idx >= 64).any()) and static is None:
            raise L.NBodyB200Error("this classifier reads static features: pass static_features")
        B = int(dyn.shape[0])
        prob = torch.empty((B,), dtype=torch.float32, device=self.device)
        label = torch.empty((B,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().nb_mlp_classify_f32(
                L.ptr(dyn), L.ptr(static), L.ptr(self.idx), self.F, L.ptr(self.mean), L.ptr(self.inv_scale), L.ptr(self.w1),
                L.ptr(self.b1), L.ptr(self.w2), L.ptr(self.b2), L.ptr(self.w3), self.b3, self.threshold, B, L.ptr(prob),
                L.ptr(label), L.stream_ptr()), "nb_mlp_classify_f32")
        return prob, label
