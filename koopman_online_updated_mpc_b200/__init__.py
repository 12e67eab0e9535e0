"""koopman_online_updated_mpc_b200 -- B200-native batched closed-loop Koopman-MPC hot path.

Python here is host glue only: every operator calls hand-written sm_100a CUDA kernels through
the C ABI of libkmpc.so (include/kmpc.h).  There is no CPU fallback.  The module layout mirrors
the reference's call sites (duffing.py / vanderpol.py / duffing_RBF.py / Tank_System.m):

    lift.Encoder(x), lift.rbf(x, cx)          stage 1  (duffing.py:764, duffing_RBF.py:20-23)
    edmd.edmd(PHIX, PHIY, U, X)               stage 2  (duffing.py:167-177)
    rls.rls_update(state, z, u, y, x_next)    stage 3  (duffing.py:927-984)
    mpc.mpc_first_move(A, B, C, z0, r, lb, ub) stage 4 (duffing.py:776-778, Tank_System.m:188)
    plant.f_update(x, u, params)              plant    (duffing.py:250-261)
    closed_loop.ClosedLoop(...).run(T)        fused scenario steps (duffing.py:823-992)
    data_generate.generate(N, N_Traj)         snapshot generator (data_generate.py:12-152)
    predict.open_loop_predict(...)            open-loop model check (duffing.py:290-343)
    io_mat.save_model_weights / save_nn_encoder / save_trajectory   the reference's .mat layouts
    losses.training_losses(enc, dec, X, U, A, B)  rec / multi-step lin / pred losses (duffing.py:179-235)
    scripts.run_duffing / run_vanderpol / run_rbf / run_tank / run_koopman_update / run_tracking_lift
                                              the reference's driver scripts, call sites swapped
"""
from . import closed_loop, data_generate, distributed, edmd, io_mat, lift, losses, mpc, plant, predict, rls, scripts, weights  # noqa: F401
from ._lib import KmpcError, launch_count, lib, measure_fp64_peak  # noqa: F401
from .build import build  # noqa: F401
from .closed_loop import ClosedLoop, LoopSpec, duffing_spec, rbf_spec, tank_spec, vanderpol_spec  # noqa: F401
from .lift import Encoder, rbf  # noqa: F401
from .rls import RLSState, rls_update  # noqa: F401

__version__ = "0.1.0"
