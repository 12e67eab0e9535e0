"""CPU oracle for the closed-loop Koopman-MPC hot path -- TEST INFRASTRUCTURE ONLY.

This package is a float64 numpy/scipy restatement of the algorithms in the reference
repository (MichaelMillerCSU/Koopman-online-updated-MPC), each function citing the reference
file:line it follows.  It exists so that the CUDA product path can be checked against the
reference's arithmetic on machines that have neither the reference nor MATLAB.

Rules (enforced by tests/test_abi.py::test_package_never_imports_the_oracle):
  * only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
    `bench.py` may import this package -- and only as the checker / the timed CPU baseline;
  * nothing under `koopman_online_updated_mpc_b200/` imports it; the product has no CPU
    fallback.

Pinning status (DESIGN.md "Oracle"):
  * Python path (duffing.py, vanderpol.py, duffing_RBF.py, vanderpol_RBF.py, data_generate.py):
    PINNED against outputs of the reference scripts run in the build container
    (tests/golden/ref_*.npz, produced by tests/golden/make_golden.py) and against the
    reference's own golden file VDP_Revise_2/NN_Encoder.mat (tests/golden/vdp_nn_encoder_head.npz).
  * MATLAB path (Tank_System.m, Koopman_update.m, Koopman_update_Tracking_Lift.m, rbf.m,
    Encoder_*.m): PARITY UNPINNED -- no MATLAB/Octave here and the reference ships no MATLAB
    outputs; restated from the sources, checked only for internal consistency (quadprog ==
    exact box-QP minimiser, RLS == batch regression identities).
"""
