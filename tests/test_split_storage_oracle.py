"""oracle/split_storage.py on the CPU: the storage format itself (32 significant bits, idempotent, ties to even) and what
it buys -- the distance from the float64 reference on a reference script stays orders of magnitude inside the 1e-5 gate
where plain float32 storage does not."""
import numpy as np

from oracle import restated as R, scenarios as SC, split_storage as SS


def test_split_round_keeps_32_bits_and_is_idempotent():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(20000) * np.exp(rng.uniform(-40, 40, 20000))
    y = SS.split_round(x)
    assert np.array_equal(SS.split_round(y), y)
    assert np.max(np.abs(y - x) / np.abs(x)) <= 2.0 ** -33 * (1 + 2.0 ** -20)
    for special in (0.0, -0.0, 1.0, -1.5, 2.0 ** -140, 3.0e38):
        assert SS.split_round(np.float64(special)) == np.float64(np.float32(special)) + SS.bf16_round(
            np.float32(special - np.float64(np.float32(special))))
    assert SS.split_round(np.array([1.0, 0.5, 1024.0]))[1] == 0.5               # fp32-representable values are kept exactly


def test_bf16_rounding_is_to_nearest_even():
    one = np.float32(1.0)
    ulp = np.float32(2.0 ** -7)                                                  # bf16 spacing at 1.0
    assert SS.bf16_round(np.array([one + ulp / 2], dtype=np.float32))[0] == one               # tie -> even (1.0)
    assert SS.bf16_round(np.array([one + ulp + ulp / 2], dtype=np.float32))[0] == one + 2 * ulp   # tie -> even (1 + 2 ulp)
    assert SS.bf16_round(np.array([one + ulp * np.float32(0.75)], dtype=np.float32))[0] == one + ulp


def test_distance_from_the_float64_reference_on_a_reference_script():
    sc = SC.spec("lens")
    f, setup, times, _ = SC.restate(sc)
    g, setup2, _, _ = SC.restate(sc)
    worst = 0.0
    SS.store(g)
    for n, t in enumerate(times[:700]):
        R.leapfrog(f, setup, t)
        SS.leapfrog(g, setup2, t)
        if (n + 1) % 50 == 0:
            worst = max(worst, np.linalg.norm(f.h - g.h) / np.linalg.norm(f.h))
    assert 0 < worst < 1e-8, worst
