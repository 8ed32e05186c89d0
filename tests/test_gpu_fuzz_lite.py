"""A fixed-seed slice of tests/fuzz/fuzz_gpu.py inside the GPU suite: random MIXED batches (transliteration / SNIPS /
cipher / random-DAG lattices of random sizes in one batch; packed jointly or part by part + concat_packed; per-arc /
theta / both score modes; integer scores in a quarter of them; state dtype auto or float64) against the C oracle at the
tests' tolerances.  Such batches found what single-shape tests had not: float32-state small-lattice posteriors at 2e-5,
float64 state forced by one deep lattice onto the float32-sized DP rings of wide shallow ones."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_random_mixed_batches_match_the_oracle(seed):
    from tests.fuzz import fuzz_gpu as F

    F.rng = np.random.default_rng(seed)
    done = 0
    while done < 25:
        names, parts = zip(*[F.part() for _ in range(int(F.rng.integers(1, 4)))])
        F.run_case(names, parts)  # raises on any mismatch or refusal
        done += 1
