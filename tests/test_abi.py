"""The C-ABI boundary: the library builds, loads, exports every symbol include/kmpc.h declares,
validates arguments, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import helpers as H
import koopman_online_updated_mpc_b200 as K
from koopman_online_updated_mpc_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(H.ROOT, "include", "kmpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmpc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 20
    handle = ctypes.CDLL(K.build())
    for name in declared:
        assert hasattr(handle, name), "declared in kmpc.h but not exported: " + name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared  # the ctypes binding covers exactly the header


def test_version_and_error_strings():
    L = K.lib()
    assert L.kmpc_version() == 100
    assert L.kmpc_strerror(0) == b"ok" and b"argument" in L.kmpc_strerror(-1)
    assert K.lib().kmpc_gram_pack_len(8, 2) == (9 + 8 + 2) * 9 + 1


def test_struct_layout_matches_header():
    """ctypes mirrors of kmpc_loop_config / kmpc_loop_buffers: field order and count."""
    text = open(os.path.join(H.ROOT, "include", "kmpc.h")).read()
    for cname, cls in (("kmpc_loop_config", _lib.LoopConfigC), ("kmpc_loop_buffers", _lib.LoopBuffersC)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if stmt:
                decl = stmt.split(",")
                names += [re.sub(r"[\s\*]", "", d).split()[-1] if False else re.findall(r"([A-Za-z_0-9]+)\s*$", d.strip())[0]
                          for d in decl]
        mine = [("lambda" if n == "lam" else n) for n, _ in cls._fields_]
        assert names == mine, (cname, names, mine)


def test_argument_validation_without_touching_the_gpu():
    L = K.lib()
    assert L.kmpc_rls_update(*([None] * 11), 4, 8, 2, 1.0, 1, None) == -1
    assert L.kmpc_qp_first_move(*([None] * 8), 100.0, 1e-4, 10, 2, 8, 4, 0, None, None, None, 0, 0.0, None) == -1
    assert L.kmpc_plant_step(None, None, None, None, 4, 0, 0, 0.05, None) == -1
    assert L.kmpc_encode(None, None, None, 4, 0, None) == -1
    a = np.zeros(4096)
    p = a.ctypes.data
    assert L.kmpc_rls_update(*([p] * 11), 4, 99, 2, 1.0, 1, None) == -1   # nz out of range
    assert L.kmpc_rls_update(*([p] * 11), 0, 8, 2, 1.0, 1, None) == 0     # empty batch is a no-op
    assert L.kmpc_qp_first_move(*([p] * 8), 100.0, 1e-4, 999, 2, 8, 4, 0, p, None, None, 0, 0.0, None) == -1
    assert L.kmpc_plant_step(p, p, p, p, 0, 0, 0, 0.05, None) == 0
    # snapshot generator / open-loop predictor / trajectory-aware Gram
    assert L.kmpc_generate_snapshots(None, None, None, 0, 0, 0.05, 4, 10, None, None, None, None) == -1
    assert L.kmpc_generate_snapshots(p, p, p, 7, 0, 0.05, 4, 10, p, p, p, None) == -1      # unknown plant
    assert L.kmpc_generate_snapshots(p, p, p, 0, 0, 0.05, 0, 10, p, p, p, None) == 0       # empty set
    assert L.kmpc_open_loop_predict(p, p, p, p, p, p, 99, 2, 1, 10, 10, 10, 0, p, p, p, None) == -1   # nz out of range
    assert L.kmpc_open_loop_predict(p, p, p, p, p, p, 8, 2, 1, 10, 0, 10, 0, p, p, p, None) == -1    # stride < 1
    assert L.kmpc_window_losses(p, p, p, p, 99, 2, 4, 31, 0, 1, p, None) == -1                       # nz out of range
    assert L.kmpc_window_losses(p, p, p, p, 8, 2, 4, 31, 0, 0, p, None) == -1                        # stride < 1
    assert L.kmpc_window_losses(p, p, p, p, 8, 2, 0, 31, 0, 1, p, None) == 0                         # no windows
    assert L.kmpc_encode_ex(None, p, p, 4, 0, 1, None) == -1
    assert L.kmpc_gram_from_snapshots_ex(None, 0, 1, p, p, p, 4, p, None) == -1
    assert L.kmpc_open_loop_predict(p, p, p, p, p, p, 8, 2, 1, 10, 10, 0, 0, p, p, p, None) == -1    # reset_every < 1
    assert L.kmpc_open_loop_predict(p, p, p, p, p, p, 8, 2, 1, 10, 10, 10, 5, p, p, p, None) == -1   # rmse row out of range
    assert L.kmpc_open_loop_predict(p, p, p, p, p, p, 8, 2, 0, 10, 10, 10, 0, p, p, p, None) == 0    # no sequences
    assert L.kmpc_gram_from_trajectories(None, 0, p, p, p, 4, 10, p, None) == -1                      # no encoder
    # closed-loop context: the explicit kernel-path / QP-start fields are range-checked before anything is allocated
    from dataclasses import replace
    import ctypes
    from koopman_online_updated_mpc_b200 import _lib
    from koopman_online_updated_mpc_b200.closed_loop import make_config, rbf_spec
    buf = _lib.LoopBuffersC(**{k: p for k in ("x", "z", "u_prev", "A", "B", "C", "KA", "P", "barX", "barQ", "r",
                                               "params_pre", "params_post", "cx")})
    for bad in (dict(qp_cold=4), dict(qp_cold=-1), dict(path=7)):
        cfg = make_config(replace(rbf_spec(), **bad), 4, False)
        ctx = ctypes.c_void_p()
        assert L.kmpc_ctx_create(ctypes.byref(ctx), ctypes.byref(cfg), ctypes.byref(buf), None, 0, None) == -1, bad


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    Ws, bs = H.oracle_weights("duffing")
    with pytest.raises(K.KmpcError):
        K.Encoder(Ws, bs)
    with pytest.raises(K.KmpcError):
        K.lift.rbf(np.zeros((4, 2)), np.zeros((8, 2)))
    with pytest.raises(K.KmpcError):
        K.mpc.mpc_first_move(np.eye(8), np.ones(8), np.ones((2, 8)), np.zeros((4, 8)), np.zeros((4, 2)), -2, 2)
    # a raw ABI compute call on host memory without a device reports a CUDA error, never computes
    a = np.zeros(64)
    rc = K.lib().kmpc_plant_step(a.ctypes.data, a.ctypes.data, a.ctypes.data, a.ctypes.data, 4, 0, 0, 0.05, None)
    assert rc == -2 and len(K.lib().kmpc_last_cuda_error()) > 0


def test_package_never_imports_the_oracle():
    pkg = os.path.join(H.ROOT, "koopman_online_updated_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "hostemu" not in text or f in ("percase.cuh", "loopbody.cuh"), f
