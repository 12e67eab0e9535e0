"""Oracle: open-loop multi-step predictor (test infrastructure -- only tests/, smoke() and bench.py's
CPU legs may import this package).

Restates duffing.py:290-343 / vanderpol.py:292-348: along the first `plot_time` snapshots of a
trajectory-major snapshot set, the lifted state is re-encoded from the TRUE state every
`reset_every` = 10 steps (`if i % 10 == 0: phix = net.Encoder(inputs_x[:, i])`) and propagated with
the EDMD model in between (`phix = A @ phix + B @ u`); the read-out `C @ phix` and the lifted state
are logged BEFORE the propagation.  RMSE as the reference defines it (duffing.py:341 row 0,
vanderpol.py:346 row 1): || (test_Y[row] - X[row, :plot_time]) / plot_time ||_2.
Pinned against the reference's own run: tests/golden/ref_{duffing,vanderpol}_predict.npz.
"""
import numpy as np


def open_loop_predict(lift_fn, A, B, C, X, U, plot_time, reset_every=10):
    """lift_fn: (n,) -> (nz,); X (n, M), U (1, M) trajectory-major snapshots.
    Returns test_Y (n, plot_time), decoder_X (nz, plot_time), marker_X (nz, ceil(plot_time/reset_every))."""
    A, B, C = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64), np.asarray(C, dtype=np.float64)
    nz = A.shape[0]
    test_Y = np.zeros((C.shape[0], plot_time))
    decoder_X = np.zeros((nz, plot_time))
    markers = []
    phix = np.asarray(lift_fn(X[:, 0]), dtype=np.float64).reshape(nz, 1)
    for i in range(plot_time):
        if i % reset_every == 0:
            phix = np.asarray(lift_fn(X[:, i]), dtype=np.float64).reshape(nz, 1)
            markers.append(phix[:, 0].copy())
        u = np.asarray(U[:, i], dtype=np.float64).reshape(1, 1)
        decoder_X[:, i] = phix[:, 0]
        test_Y[:, i] = (C @ phix)[:, 0]
        phix = A @ phix + B @ u
    return test_Y, decoder_X, np.stack(markers, axis=1)


def rmse(test_Y, X, plot_time, row=0):
    """duffing.py:341 (row 0) / vanderpol.py:346 (row 1)."""
    return float(np.linalg.norm((test_Y[row, :] - X[row, :plot_time]) / plot_time, ord=2))
