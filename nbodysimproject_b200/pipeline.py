"""MLTrainingPipeline with the reference's entry points (ml_training_pipeline.py:27-237).

Systems are generated on the host with the same global-RNG draw order, constructed as NBodySimulation objects
(default integrator mode, i.e. ham_soft, like the reference) and analysed in ONE batched GPU pass per bucket
instead of a Python loop over systems.  `integrator_mode=` is an extension that lets a caller pick the
classic integrators (BASELINE.json's C2/C3 wording: verlet / yoshida4 with MEGNO)."""
from __future__ import annotations

import numpy as np

from .generators import GeneratorConfig, InitialConditionGenerator, SpecializedGenerators, set_global_seed
from .simulation import NBodySimulation
from .stability import BatchStabilityAnalyzer, StabilityAnalyzer, analyze_simulations


class MLTrainingPipeline:
    def __init__(self, n_systems: int = 1000, n_steps: int = 1000, dt: float = 0.01, integrator_mode: str | None = None):
        self.n_systems = n_systems
        self.n_steps = max(500, min(2000, n_steps))
        self.dt = dt
        self.integrator_mode = integrator_mode
        self.ic_generator = InitialConditionGenerator()
        self.batch_analyzer = BatchStabilityAnalyzer(n_steps=self.n_steps, dt=self.dt, mode="full")

    def _kw(self):
        return {} if self.integrator_mode is None else {"integrator_mode": self.integrator_mode}

    def generate_diverse_dataset(self):
        """ml_training_pipeline.py:39-135: 40 % random, 30 % hierarchical, 20 % polygons, 10 % close encounters."""
        print(f"Generating {self.n_systems} diverse N-body systems...")
        sims = []
        n_random = int(0.4 * self.n_systems)
        print(f"\n1. Generating {n_random} random systems...")
        for i in range(n_random):
            n_bodies = np.random.randint(3, 6)
            cfg = GeneratorConfig(mass_range=(0.1, 10.0), use_log_mass=(i % 2 == 0),
                                  position_scale=np.random.uniform(0.5, 2.0),
                                  velocity_virial_fraction=np.random.uniform(0.8, 1.2),
                                  velocity_perturbation=np.random.uniform(0.05, 0.2),
                                  softening=np.random.uniform(0.001, 0.1))
            sims.append(InitialConditionGenerator(cfg).create_simulation(n_bodies, **self._kw()))
        n_hier = int(0.3 * self.n_systems)
        print(f"2. Generating {n_hier} hierarchical systems...")
        for i in range(n_hier):
            r1 = np.random.uniform(0.1, 1.0)
            r2 = np.random.uniform(0.1, 2.0)
            sep = np.random.uniform(3, 50)
            m, p, v = SpecializedGenerators.generate_hierarchical_triple(r1, r2, sep)
            v += np.random.randn(*v.shape) * 0.05
            sims.append(NBodySimulation(masses=m, positions=p, velocities=v, G=1.0, softening=0.01, **self._kw()))
        n_poly = int(0.2 * self.n_systems)
        print(f"3. Generating {n_poly} polygon configurations...")
        for i in range(n_poly):
            n_bodies = np.random.randint(3, 8)
            radius = np.random.uniform(0.5, 3.0)
            rot = np.random.uniform(0, 1.0)
            m, p, v = SpecializedGenerators.generate_equal_mass_polygon(n_bodies, radius, rot)
            sims.append(NBodySimulation(masses=m, positions=p, velocities=v, G=1.0, softening=0.05, **self._kw()))
        n_close = self.n_systems - n_random - n_hier - n_poly
        print(f"4. Generating {n_close} close encounter systems...")
        for i in range(n_close):
            n_bodies = np.random.randint(3, 5)
            cfg = GeneratorConfig(position_scale=0.1, velocity_virial_fraction=1.5, velocity_perturbation=0.3,
                                  softening=0.001)
            sims.append(InitialConditionGenerator(cfg).create_simulation(n_bodies, **self._kw()))
        print(f"\nAnalyzing {len(sims)} systems...")
        df = self.batch_analyzer.analyze_batch(sims, show_progress=True)
        df["system_type"] = (["random"] * n_random + ["hierarchical"] * n_hier + ["polygon"] * n_poly
                             + ["close_encounter"] * n_close)
        return df

    def generate_focused_dataset(self, focus: str = "boundary"):
        """ml_training_pipeline.py:137-199."""
        print(f"Generating {self.n_systems} systems focused on {focus} cases...")
        sims = []
        if focus == "boundary":
            for i in range(self.n_systems):
                if i % 3 == 0:
                    sep = np.random.uniform(5, 15)
                    m, p, v = SpecializedGenerators.generate_hierarchical_triple(separation_ratio=sep)
                    sim = NBodySimulation(masses=m, positions=p, velocities=v, **self._kw())
                elif i % 3 == 1:
                    cfg = GeneratorConfig(velocity_virial_fraction=1.0, velocity_perturbation=np.random.uniform(0.1, 0.3))
                    gen = InitialConditionGenerator(cfg)
                    sim = gen.create_simulation(np.random.randint(3, 5), **self._kw())
                else:
                    n = np.random.randint(4, 7)
                    rot = np.random.uniform(0.3, 0.7)
                    m, p, v = SpecializedGenerators.generate_equal_mass_polygon(n, rotation_fraction=rot)
                    sim = NBodySimulation(masses=m, positions=p, velocities=v, **self._kw())
                sims.append(sim)
        elif focus == "stable":
            for i in range(self.n_systems):
                sep = np.random.uniform(20, 100)
                m, p, v = SpecializedGenerators.generate_hierarchical_triple(separation_ratio=sep)
                v += np.random.randn(*v.shape) * 0.01
                sims.append(NBodySimulation(masses=m, positions=p, velocities=v, softening=0.01, **self._kw()))
        else:
            for i in range(self.n_systems):
                cfg = GeneratorConfig(position_scale=0.1, velocity_virial_fraction=np.random.uniform(1.5, 2.0),
                                      velocity_perturbation=0.5, softening=0.001)
                gen = InitialConditionGenerator(cfg)
                sims.append(gen.create_simulation(np.random.randint(3, 6), **self._kw()))
        df = self.batch_analyzer.analyze_batch(sims)
        df["dataset_focus"] = focus
        return df

    def quick_test_pipeline(self):
        """ml_training_pipeline.py:201-235: 10 systems, 100 steps, 'core' mode.  Divergent ham_soft systems come
        back with inf/NaN drifts and a status flag instead of the reference's OverflowError."""
        import pandas as pd
        set_global_seed(42)
        print("Running quick test with 10 systems...")
        gen = InitialConditionGenerator()
        sims = [gen.create_simulation(3 + (i % 3), **self._kw()) for i in range(10)]
        print("\nTesting unified analyzer in core mode...")
        rows = analyze_simulations(sims, 100, 0.01, "core")
        results = []
        for i, row in enumerate(rows):
            for k in ("pathological_energy", "softening_policy", "_status"):
                row.pop(k, None)
            row["system_id"] = i
            results.append(row)
            print(f"System {i}: {'STABLE' if row['is_stable'] else 'UNSTABLE'} (E_drift={row['energy_drift']:.2e})")
        df = pd.DataFrame(results)
        n_stable = int(sum(df["is_stable"]))
        print(f"\nTest complete. {n_stable} stable, {len(df) - n_stable} unstable")
        return df
