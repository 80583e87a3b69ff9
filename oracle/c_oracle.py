"""TEST INFRASTRUCTURE ONLY: builds and binds oracle/vq_oracle.c (plain-C restatement of the index-producing tail of the encode
path and of the DiTi token schedule).  Imported by tests/ only; the product package never imports anything under oracle/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "vq_oracle.c")
OUT_DIR = os.path.join(_HERE, "_build")
LIB = os.path.join(OUT_DIR, "libselftok_oracle_c.so")


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off: no implicit contraction -- the only fused operations are the explicit fmaf() calls."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on oracle/vq_oracle.c:\n" + r.stderr)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for fn in ("selftok_oracle_vq", "selftok_oracle_lookup_ln", "selftok_oracle_diti_k"):
            getattr(_lib, fn).restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def vq(z: np.ndarray, proj_w: np.ndarray, proj_b, embed: np.ndarray):
    """z [n, qdim] fp32 -> (ids [n] int64, x_hat [n, cdim], margin [n])"""
    z = np.ascontiguousarray(z, np.float32)
    proj_w = np.ascontiguousarray(proj_w, np.float32)
    embed = np.ascontiguousarray(embed, np.float32)
    pb = np.ascontiguousarray(proj_b, np.float32) if proj_b is not None else None
    n, qdim = z.shape
    ncode, cdim = embed.shape
    ids = np.empty(n, np.int64)
    xh = np.empty((n, cdim), np.float32)
    mg = np.empty(n, np.float32)
    st = lib().selftok_oracle_vq(_p(z), C.c_int64(n), qdim, _p(proj_w), _p(pb) if pb is not None else None, _p(embed), ncode, cdim,
                                 _p(ids), _p(xh), _p(mg))
    if st:
        raise RuntimeError(f"selftok_oracle_vq -> {st}")
    return ids, xh, mg


def lookup_ln(ids: np.ndarray, embed: np.ndarray, ln_w: np.ndarray, ln_b: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    ids = np.ascontiguousarray(ids, np.int64).reshape(-1)
    embed = np.ascontiguousarray(embed, np.float32)
    out = np.empty((ids.shape[0], embed.shape[1]), np.float32)
    st = lib().selftok_oracle_lookup_ln(_p(ids), C.c_int64(ids.shape[0]), _p(embed), embed.shape[0], embed.shape[1],
                                        _p(np.ascontiguousarray(ln_w, np.float32)), _p(np.ascontiguousarray(ln_b, np.float32)),
                                        C.c_float(eps), _p(out))
    if st:
        raise RuntimeError(f"selftok_oracle_lookup_ln -> {st}")
    return out


def diti_k(t_mapped: np.ndarray, stages, k_per_stage, K: int) -> np.ndarray:
    t = np.ascontiguousarray(t_mapped, np.int64)
    st_ = np.ascontiguousarray(stages, np.int32)
    kp = np.ascontiguousarray(k_per_stage, np.int32)
    out = np.empty(t.shape[0], np.int64)
    st = lib().selftok_oracle_diti_k(_p(t), int(t.shape[0]), _p(st_), _p(kp), int(kp.shape[0]), int(K), _p(out))
    if st:
        raise RuntimeError(f"selftok_oracle_diti_k -> {st}")
    return out
