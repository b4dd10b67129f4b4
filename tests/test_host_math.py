"""Host-side mirrors in libmmo_b200.so (libm, IEEE double) against the oracle: bit-identical."""
import numpy as np

import mmo_b200


def test_so3_rotations_bit_identical(orc):
    for n in (1, 4, 1000):
        assert np.array_equal(mmo_b200.SO3.rotations(n), orc.so3_rotations(n))


def test_rxyz_and_decompose_bit_identical(orc):
    rng = np.random.default_rng(0)
    for _ in range(100):
        a, b, g = rng.uniform(-3, 3), rng.uniform(-1.5, 1.5), rng.uniform(-3, 3)
        r = mmo_b200.Rot.r_xyz(a, b, g)
        assert np.array_equal(r, orc.rot_r_xyz(a, b, g))
        assert np.array_equal(mmo_b200.Rot.decompose(r), orc.rot_decompose(r))


def test_grid_from_box(orc):
    for step, box in ((0.5, (108.76, 101.3, 116.9)), (0.375, (30.0, 30.0, 30.0)), (1.0, (20.0, 20.0, 20.0)),
                      (2.0, (20.0, 20.0, 20.0)), (0.5, (20.0, 20.0, 20.0))):
        assert mmo_b200.Grid.from_box(step, *box) == orc.grid_from_box(step, *box)
    assert mmo_b200.Grid.from_box(0.375, 30.0, 30.0, 30.0) == (81, 81, 81)     # config C3
    assert mmo_b200.Grid.from_box(1.0, 20.0, 20.0, 20.0) == (21, 21, 21)       # config C2 lattice
