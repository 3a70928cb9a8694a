"""Device logic of the valley / PMD split steps (tests/emu/emu_valleys.cpp) against the restatements on tables of very
different shapes: tiny conditions, all-low and all-high tables, one-base segments that touch, large gaps, weights over eight
decades, extreme thresholds.  (1200 / 600 trials of the same generator passed when the steps were written.)"""
import numpy as np
import test_valleys as T
from test_valleys import emu  # noqa: F401


def rand_table(rng):
    mode = rng.integers(0, 6)
    beta, width, gap, ncpg, cond, job, pw, cov = [], [], [], [], [], [], [], []
    for j in range(int(rng.integers(1, 4))):
        for c in range(1, int(rng.integers(2, 4))):
            n = int(rng.integers(1, 8)) if mode == 0 else int(rng.integers(1, 200))
            plow = [0.55, 1.0, 0.0, 0.9, 0.1, 0.5][mode]
            low = rng.random(n) < plow
            beta += list(np.where(low, rng.random(n) * 0.12, 0.2 + rng.random(n) * 0.8))
            width += list(rng.integers(1, 6 if mode == 4 else 9000, n).astype(float))
            gap += list(np.where(rng.random(n) < 0.15, rng.integers(4000, 12000, n), rng.integers(0, 3 if mode in (3, 4) else 2000, n)).astype(float))
            ncpg += list(rng.integers(1, 80, n)); cond += [c] * n; job += [j] * n
            pw += list(rng.random(n) * 10.0 ** rng.integers(-2, 6)); cov += list(rng.integers(5, 5000, n).astype(float))
    return T.table(beta, width, gap, ncpg, cond, job, pw, cov)


def test_valleys_on_diverse_tables(emu, oracle):
    rng = np.random.default_rng(2024)
    tot = 0
    for trial in range(240):
        t = rand_table(rng)
        n = t["start"].size
        category = np.where(rng.random(n) < 0.25, 2, np.where(rng.random(n) < 0.2, 1, 0)).astype(np.int8)
        kw = [{}, dict(valley_max_beta=0.05, pmd_valley_min_width=2500.0), dict(pmd_valley_min_width=800.0, max_distance=200.0, min_num_cpgs=20),
              dict(valley_max_beta=0.2, pmd_valley_min_width=10.0)][trial % 4]
        got = emu(t, category, **kw)
        tot += T.check_against_oracle(got, t, category, **kw)
    assert tot > 5000


def test_pmd_split_on_diverse_tables(emu, oracle):
    rng = np.random.default_rng(77)
    tot = nv = 0
    for trial in range(120):
        a, b = T.split_trial(emu, rng, rand_table(rng), trial)
        tot += a; nv += b
    assert tot > 5000 and nv > 1000
