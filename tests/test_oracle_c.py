"""CPU: the C/OpenMP oracle (oracle/pencil_c.c) equals the NumPy pencil oracle bit for bit."""
import numpy as np

from oracle import pencil, pencil_c


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(3)
    f = rng.standard_normal((13, 17, 22))
    for axis in range(3):
        for order in (1, 2):
            for acc in (2, 4, 6):
                assert np.array_equal(pencil_c.fd_apply(f, axis, 0.037, order, acc),
                                      pencil.fd_apply(f, axis, 0.037, order, acc))
        n = f.shape[axis]
        M = pencil.chebyshev_matrix(n, 0.0, 1.5, 2, last_axis_of_nd=(axis == 2))
        for desc in (False, True):
            assert np.array_equal(pencil_c.dense_apply(f, axis, M, desc),
                                  pencil.chebyshev_apply(f, axis, M, desc))
    x = np.logspace(-1, 1, 17)
    assert np.array_equal(pencil_c.fd_apply(f, 1, 0.05, 2, 6, x_div=x), pencil.log_apply(f, 1, 0.05, x, 2, 6))
    hs = (0.1, 0.05, 0.025)
    for acc in (2, 4, 6):
        assert np.array_equal(pencil_c.cartesian_laplacian3(f, hs, acc), pencil.cartesian_laplacian(f, hs, acc))
    assert pencil_c.threads() >= 1
