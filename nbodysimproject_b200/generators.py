"""Initial-condition generators (host side, NumPy global RNG).

Mirrors initial_condition_generator.py:30-170 and specialized_generators.py:23-94 of the reference with the
SAME sequence of global-RNG draws, so a seeded run produces bit-identical systems; these are the synthetic
inputs of every benchmark configuration.  `EnsembleInputs` adds vectorised cohort
generators for million-system batches, where constructing Python objects one by one would dominate.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np


def set_global_seed(seed: int):
    """utils.py:17-28 (torch seeding included so RNG streams line up with the reference)."""
    random.seed(seed)
    np.random.seed(seed)
    try:
        import torch
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
    except Exception:  # pragma: no cover - torch is always present in this image
        pass
    print(f"Global random seed set to {seed}")


def remove_center_of_mass_velocity(masses: np.ndarray, velocities: np.ndarray) -> np.ndarray:
    """physics_utils.py:16-26."""
    if len(masses) == 1:
        return velocities.copy()
    total = float(np.sum(masses))
    if total == 0 or velocities.size == 0:
        return velocities.copy()
    return velocities - np.sum(masses[:, None] * velocities, axis=0) / total


@dataclass
class GeneratorConfig:
    mass_range: Tuple[float, float] = (0.1, 10.0)
    use_log_mass: bool = False
    position_scale: float = 1.0
    velocity_virial_fraction: float = 1.0
    velocity_perturbation: float = 0.1
    softening: float = 0.05
    G: float = 1.0
    seed: Optional[int] = None


class InitialConditionGenerator:
    """initial_condition_generator.py:41-170."""

    def __init__(self, config: GeneratorConfig | None = None):
        self.config = config or GeneratorConfig()
        if self.config.seed is not None:
            np.random.seed(self.config.seed)

    def _generate_masses(self, n: int) -> np.ndarray:
        lo, hi = self.config.mass_range
        if self.config.use_log_mass:
            return np.exp(np.random.uniform(np.log(lo), np.log(hi), n))
        return np.random.uniform(lo, hi, n)

    def _generate_positions(self, n: int) -> np.ndarray:
        return np.random.randn(n, 2) * self.config.position_scale

    @staticmethod
    def _compute_mean_separation(positions: np.ndarray) -> float:
        n = len(positions)
        if n < 2:
            return 1.0
        d = positions[:, None, :] - positions[None, :, :]
        dist = np.sqrt((d ** 2).sum(axis=-1))
        iu = np.triu_indices(n, 1)
        return float(np.mean(dist[iu])) if iu[0].size else 1.0

    def _compute_potential_energy(self, m: np.ndarray, pos: np.ndarray) -> float:
        # note: additive softening r + eps here, not Plummer (initial_condition_generator.py:72-80)
        G, eps = self.config.G, self.config.softening
        U = 0.0
        for i in range(len(m) - 1):
            for j in range(i + 1, len(m)):
                U -= G * m[i] * m[j] / (np.hypot(*(pos[j] - pos[i])) + eps)
        return U

    def _generate_velocities(self, m: np.ndarray, pos: np.ndarray) -> np.ndarray:
        n, G = len(m), self.config.G
        K_target = -self._compute_potential_energy(m, pos) / 2.0 * self.config.velocity_virial_fraction
        if K_target <= 0.0:
            v_char = np.sqrt(G * m.sum() / self._compute_mean_separation(pos))
        else:
            v_char = np.sqrt(2.0 * K_target / m.sum())
        vel = np.random.randn(n, 2)
        speed = np.linalg.norm(vel, axis=1, keepdims=True)
        vel = np.where(speed > 0, vel / speed * v_char, vel)
        vel = remove_center_of_mass_velocity(m, vel)
        vel += np.random.randn(n, 2) * v_char * self.config.velocity_perturbation
        return remove_center_of_mass_velocity(m, vel)

    def generate_single(self, n_bodies: int):
        m = self._generate_masses(n_bodies)
        p = self._generate_positions(n_bodies)
        v = self._generate_velocities(m, p)
        return m, p, v

    def generate_batch(self, n_systems: int, n_bodies_range: Tuple[int, int] = (3, 5)):
        out = []
        for _ in range(n_systems):
            n = np.random.randint(n_bodies_range[0], n_bodies_range[1] + 1)
            out.append(self.generate_single(n))
        return out

    def create_simulation(self, n_bodies: int, *, integrator_mode: str | None = None,
                          adaptive_softening: bool | None = None):
        from .simulation import NBodySimulation
        m, p, v = self.generate_single(n_bodies)
        kw: Dict = dict(masses=m, positions=p, velocities=v, G=self.config.G, softening=self.config.softening)
        if integrator_mode is not None:
            kw["integrator_mode"] = integrator_mode
        if adaptive_softening is not None:
            kw["adaptive_softening"] = adaptive_softening
        return NBodySimulation(**kw)

    def validate_system(self, masses, positions, velocities) -> Dict[str, float]:
        from .simulation import NBodySimulation
        from .stability import Diagnostics
        sim = NBodySimulation(masses=masses, positions=positions, velocities=velocities, G=self.config.G,
                              softening=self.config.softening)
        d = Diagnostics(sim)
        KE, PE = d.kinetic_energy(), d.potential_energy()
        com_pos, com_vel = d.center_of_mass()
        return {
            "kinetic_energy": KE, "potential_energy": PE, "total_energy": KE + PE,
            "virial_ratio": 2 * KE / abs(PE) if PE else np.inf, "angular_momentum": d.angular_momentum(),
            "com_position": float(np.linalg.norm(com_pos)), "com_velocity": float(np.linalg.norm(com_vel)),
            "is_bound": bool(KE + PE < 0),
        }


class SpecializedGenerators:
    """specialized_generators.py:21-94."""

    @staticmethod
    def generate_hierarchical_triple(mass_ratio1: float = 1.0, mass_ratio2: float = 0.5,
                                     separation_ratio: float = 10.0, G: float = 1.0, *, integrator_mode=None,
                                     adaptive_softening=None):
        m1, m2, m3 = 1.0, mass_ratio1, mass_ratio2
        masses = np.array([m1, m2, m3])
        a_in = 1.0
        a_out = max(separation_ratio * a_in, 5.0 * a_in)
        positions = np.array([[-m2 * a_in / (m1 + m2), 0.0], [m1 * a_in / (m1 + m2), 0.0], [a_out, 0.0]])
        v_in = np.sqrt(G * (m1 + m2) / a_in)
        v_out = np.sqrt(G * (m1 + m2 + m3) / a_out)
        velocities = np.array([[0.0, -m2 * v_in / (m1 + m2)], [0.0, m1 * v_in / (m1 + m2)], [0.0, v_out]])
        return masses, positions, remove_center_of_mass_velocity(masses, velocities)

    @staticmethod
    def generate_equal_mass_polygon(n_bodies: int, radius: float = 1.0, rotation_fraction: float = 0.5,
                                    G: float = 1.0, *, integrator_mode=None, adaptive_softening=None):
        masses = np.ones(n_bodies)
        ang = np.linspace(0.0, 2.0 * np.pi, n_bodies, endpoint=False)
        positions = np.column_stack([radius * np.cos(ang), radius * np.sin(ang)])
        v_scale = np.sqrt(G * float(np.sum(masses)) / radius) * rotation_fraction
        velocities = np.column_stack([-v_scale * np.sin(ang), v_scale * np.cos(ang)])
        return masses, positions, remove_center_of_mass_velocity(masses, velocities)


# ---------------------------------------------------------------------------------------------
# vectorised cohort generators for large synthetic ensembles (benchmarks)
# ---------------------------------------------------------------------------------------------

class EnsembleInputs:
    """Same distributions as ml_training_pipeline.py:44-122 (extended to N <= 8), drawn with a private
    Generator and vectorised over the batch.  Returns (m[B,N], q[B,N,2], v[B,N,2], softening[B])."""

    @staticmethod
    def _virial_velocities(rng, m, q, soft, frac, pert):
        B, N = m.shape
        d = q[:, :, None, :] - q[:, None, :, :]
        r = np.sqrt((d ** 2).sum(-1))
        iu = np.triu_indices(N, 1)
        U = -np.sum(m[:, iu[0]] * m[:, iu[1]] / (r[:, iu[0], iu[1]] + soft[:, None]), axis=1)
        K = -U / 2.0 * frac
        v_char = np.sqrt(2.0 * np.maximum(K, 1e-300) / m.sum(1))
        v = rng.standard_normal((B, N, 2))
        v = v / np.linalg.norm(v, axis=2, keepdims=True) * v_char[:, None, None]
        v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
        v += rng.standard_normal((B, N, 2)) * (v_char * pert)[:, None, None]
        v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
        return v

    @classmethod
    def random(cls, rng, B, N, close_encounter=False):
        if close_encounter:
            scale = np.full(B, 0.1); frac = np.full(B, 1.5); pert = np.full(B, 0.3); soft = np.full(B, 0.001)
        else:
            scale = rng.uniform(0.5, 2.0, B); frac = rng.uniform(0.8, 1.2, B)
            pert = rng.uniform(0.05, 0.2, B); soft = rng.uniform(0.001, 0.1, B)
        m = rng.uniform(0.1, 10.0, (B, N))
        log = np.arange(B) % 2 == 0
        if not close_encounter:
            m[log] = np.exp(rng.uniform(np.log(0.1), np.log(10.0), (int(log.sum()), N)))
        q = rng.standard_normal((B, N, 2)) * scale[:, None, None]
        return m, q, cls._virial_velocities(rng, m, q, soft, frac, pert), soft

    @staticmethod
    def hierarchical(rng, B):
        m2 = rng.uniform(0.1, 1.0, B); m3 = rng.uniform(0.1, 2.0, B); sep = rng.uniform(3, 50, B)
        m = np.stack([np.ones(B), m2, m3], 1)
        a_out = np.maximum(sep, 5.0)
        q = np.zeros((B, 3, 2))
        q[:, 0, 0] = -m2 / (1 + m2); q[:, 1, 0] = 1 / (1 + m2); q[:, 2, 0] = a_out
        v_in = np.sqrt(1 + m2); v_out = np.sqrt((1 + m2 + m3) / a_out)
        v = np.zeros((B, 3, 2))
        v[:, 0, 1] = -m2 * v_in / (1 + m2); v[:, 1, 1] = v_in / (1 + m2); v[:, 2, 1] = v_out
        v -= (m[:, :, None] * v).sum(1, keepdims=True) / m.sum(1)[:, None, None]
        v += rng.standard_normal((B, 3, 2)) * 0.05
        return m, q, v, np.full(B, 0.01)

    @staticmethod
    def polygon(rng, B, N):
        radius = rng.uniform(0.5, 3.0, B); rot = rng.uniform(0, 1.0, B)
        ang = np.linspace(0.0, 2.0 * np.pi, N, endpoint=False)
        m = np.ones((B, N))
        q = np.stack([radius[:, None] * np.cos(ang), radius[:, None] * np.sin(ang)], -1)
        vs = np.sqrt(N / radius) * rot
        v = np.stack([-vs[:, None] * np.sin(ang), vs[:, None] * np.cos(ang)], -1)
        return m, q, v, np.full(B, 0.05)

    @classmethod
    def diverse(cls, rng, B, n_max=8):
        """40 % random N in 3..n_max, 30 % hierarchical triples, 20 % polygons N in 3..7, 10 % close encounters.
        Returns {N: (m, q, v, soft, cohort_id)} buckets."""
        n_rand, n_hier, n_poly = int(0.4 * B), int(0.3 * B), int(0.2 * B)
        n_close = B - n_rand - n_hier - n_poly
        out: Dict[int, list] = {}

        def add(N, tup, cohort):
            out.setdefault(N, []).append(tup + (np.full(tup[0].shape[0], cohort, dtype=np.int8),))

        Ns = list(range(3, n_max + 1))
        per = np.bincount(rng.integers(0, len(Ns), n_rand), minlength=len(Ns))
        for N, c in zip(Ns, per):
            if c:
                add(N, cls.random(rng, int(c), N), 0)
        add(3, cls.hierarchical(rng, n_hier), 1)
        Np = list(range(3, 8))
        per = np.bincount(rng.integers(0, len(Np), n_poly), minlength=len(Np))
        for N, c in zip(Np, per):
            if c:
                add(N, cls.polygon(rng, int(c), N), 2)
        per = np.bincount(rng.integers(0, 2, n_close), minlength=2)
        for N, c in zip((3, 4), per):
            if c:
                add(N, cls.random(rng, int(c), N, close_encounter=True), 3)
        return {N: tuple(np.concatenate([t[k] for t in lst]) for k in range(5)) for N, lst in out.items()}

    @staticmethod
    def planetary(rng, B, n_planets, ttv=False):
        """C4 cohort (no generator exists in the reference; defined in SURVEY.md section 8d): star m0 = 1,
        planets log-U(1e-6, 1e-3), a1 = 1, period ratios near {3:2, 2:1, 5:3}, circular coplanar."""
        N = n_planets + 1
        m = np.ones((B, N))
        m[:, 1:] = 10 ** rng.uniform(-6, -3, (B, n_planets))
        if ttv:
            m[:, 1] *= 10
        ratios = np.array([1.5, 2.0, 5.0 / 3.0])[rng.integers(0, 3, (B, max(n_planets - 1, 1)))]
        ratios = ratios * (1 + rng.uniform(-0.02, 0.02, ratios.shape))
        P = np.cumprod(np.concatenate([np.ones((B, 1)), ratios[:, :n_planets - 1]], 1), 1)
        a = P ** (2.0 / 3.0)
        ph = rng.uniform(0, 2 * np.pi, (B, n_planets))
        q = np.zeros((B, N, 2)); v = np.zeros((B, N, 2))
        q[:, 1:, 0] = a * np.cos(ph); q[:, 1:, 1] = a * np.sin(ph)
        vc = np.sqrt((1.0 + m[:, 1:]) / a)
        v[:, 1:, 0] = -vc * np.sin(ph); v[:, 1:, 1] = vc * np.cos(ph)
        return m, q, v, np.zeros(B)


COHORTS = {"random": 0, "hierarchical": 1, "polygon": 2, "close": 3, "planetary": 4, "planetary_ttv": 5}


def generate_on_device(cohort, N: int, B: int, seed: int = 42, first_index: int = 0, device=None):
    """B systems of `cohort` ('random', 'hierarchical', 'polygon', 'close', 'planetary', 'planetary_ttv') built on
    the GPU by `nb_generate_ensemble_f64` (counter-based RNG: system k depends only on (seed, first_index + k)).
    Returns device tensors m[B,N], q[B,N,2], v[B,N,2], softening[B] ready for ensemble.DeviceBucket."""
    from . import _lib as L
    torch = L.require_cuda()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    c = COHORTS[cohort] if isinstance(cohort, str) else int(cohort)
    m = torch.empty((B, N), dtype=torch.float64, device=dev)
    q = torch.empty((B, N, 2), dtype=torch.float64, device=dev)
    v = torch.empty((B, N, 2), dtype=torch.float64, device=dev)
    eps = torch.empty((B,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().nb_generate_ensemble_f64(c, int(N), int(B), int(seed), int(first_index), L.ptr(m), L.ptr(q),
                                                  L.ptr(v), L.ptr(eps), L.stream_ptr()), "nb_generate_ensemble_f64")
    return m, q, v, eps
