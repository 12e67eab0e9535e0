"""Plant step (the environment of the closed loop).  Drop-in for `f_update(0, x, u)`
(duffing.py:250-261), the MATLAB RK4 with the k4 = f(x + k1 dt) quirk (Koopman_update.m:21-25)
and the cascaded-tank map (Tank_System.m:9-10, 211)."""
import torch

from . import _lib
from ._tensors import like_input, ptr, stream_ptr, to_dev

PLANT_POLY2, PLANT_TANK = 0, 1
RK4_PYTHON, RK4_MATLAB = 0, 1

# x1' = p0 x2 ; x2' = p1 x2 + p2 x1 + p3 x1^3 + p4 x1^2 x2 + u
DUFFING_PRE = (1.0, -0.5, 1.0, -1.0, 0.0)      # duffing.py:255
DUFFING_POST = (1.0, -5.0, 2.0, -0.5, 0.0)     # duffing.py:803
VDP_PRE = (2.0, 2.0, -0.8, 0.0, -10.0)         # vanderpol.py:252
VDP_POST = (1.0, -3.0, -3.0, 0.0, -10.0)       # vanderpol.py:714
# x1+ = x1 - p0 sqrt(x1) + p1 u ; x2+ = x2 + p2 sqrt(x1) - p3 sqrt(x2)
TANK_PRE = (0.5, 0.4, 0.2, 0.3, 0.0)           # Tank_System.m:9-10
TANK_POST = (0.53, 0.3, 0.1, 0.35, 0.0)        # Tank_System.m:194-195


def f_update(x, u, params, kind=PLANT_POLY2, rk4_variant=RK4_PYTHON, h=0.05):
    """x (S,2), u (S,) or (S,1), params (5,) shared or (S,5) -> x_next (S,2)."""
    x_d = to_dev(x).reshape(-1, 2)
    S = x_d.shape[0]
    u_d = to_dev(u).reshape(-1)
    p_d = to_dev(params)
    if p_d.ndim == 1:
        p_d = p_d.reshape(1, 5).expand(S, 5).contiguous()
    out = torch.empty_like(x_d)
    _lib.check(_lib.lib().kmpc_plant_step(ptr(x_d), ptr(u_d), ptr(p_d), ptr(out), S, kind, rk4_variant,
                                          float(h), stream_ptr()))
    return like_input(out, x)
