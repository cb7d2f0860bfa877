"""nbodysimproject_b200 -- B200-native (sm_100a) hot path of calkan27/NBodySimProject.

Public surface mirrors the reference's Python API (`minbody/__init__.py:15-129`) for the path in
BASELINE.json: NBodySimulation, the pair-kernel functions, StabilityAnalyzer / BatchStabilityAnalyzer,
the initial-condition generators and MLTrainingPipeline.  All arithmetic runs in hand-written CUDA
kernels behind the C ABI in include/nbody_b200.h; there is no CPU fallback.
"""
from . import _lib
from ._lib import NBodyB200Error

__all__ = ["_lib", "NBodyB200Error"]


def __getattr__(name):
    # lazy re-exports so that `import nbodysimproject_b200` works without torch/pandas start-up cost
    import importlib
    table = {
        "NBodySimulation": "simulation", "SimConfig": "simulation", "Body": "simulation",
        "gravitational_force": "forces", "pairwise_force": "forces", "softened_forces": "forces",
        "dV_d_epsilon": "forces", "softened_potential": "forces", "dU_d_eps": "forces", "TangentMap": "forces",
        "StabilityAnalyzer": "stability", "BatchStabilityAnalyzer": "stability", "Diagnostics": "stability",
        "DynamicalFeatures": "stability", "EvolutionFeatures": "stability",
        "InitialConditionGenerator": "generators", "GeneratorConfig": "generators",
        "SpecializedGenerators": "generators", "set_global_seed": "generators",
        "MLTrainingPipeline": "pipeline",
        "LargeNSimulation": "largen", "LargeNHamSoftSimulation": "largen",
        "StabilityClassifier": "classifier", "StabilityDataset": "dataset", "save_feature_table": "dataset", "table_from_tensors": "dataset",
    }
    if name in table:
        mod = importlib.import_module("." + table[name], __name__)
        return getattr(mod, name)
    raise AttributeError(name)
