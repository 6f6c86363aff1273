"""Seeded random gravity configurations shared by the oracle-vs-reference sweep (CPU) and the GPU-vs-oracle sweep."""
import numpy as np

from gasoline_b200 import ics


def random_case(seed):
    """A seeded random configuration: particle count, distribution (clustered / uniform in a periodic box / with exact
    duplicates), masses and softenings over two decades (large ones make softened cells), active fraction, nBucket,
    theta, list and Ewald orders, replicas."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 2, 7, 9, 40, 300, 1200, 2500]))
    periodic = bool(rng.integers(0, 2)) and n >= 40
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    if periodic:
        pos = rng.uniform(-0.5, 0.5, (n, 3)) * 0.999
        period = (1.0, 1.0, 1.0)
    else:
        pos = rng.normal(0, 1, (n, 3)) * rng.choice([0.05, 1.0, 30.0])
        period = (ics.FLOAT_MAXVAL,) * 3
    if n >= 40 and rng.integers(0, 3) == 0:  # exact duplicates (zero-extent cells)
        k = n // 10
        pos[rng.integers(0, n, k)] = pos[rng.integers(0, n, k)]
    m = rng.uniform(0.1, 10.0, n) / n
    h = 10.0 ** rng.uniform(-3.5, -1.0, n) * (0.2 if periodic else 1.0)
    p = ics.Particles(f32(pos[:, 0]), f32(pos[:, 1]), f32(pos[:, 2]), m, h, period, name=f"random{seed}")
    active = None
    if n >= 7 and rng.integers(0, 2):
        active = (rng.uniform(0, 1, n) < rng.choice([0.05, 0.5, 0.9])).astype(np.int32)
        if active.sum() == 0:
            active[0] = 1
    kw = dict(nReps=int(rng.integers(1, 3)) if periodic else 0, bPeriodic=int(periodic),
              bEwald=int(periodic and rng.integers(0, 2)), iOrder=int(rng.integers(1, 5)), iEwOrder=int(rng.integers(2, 5)))
    return p, active, int(rng.choice([1, 2, 5, 8, 16, 33])), float(rng.uniform(0.3, 1.0)), kw
