"""ctypes loader for oracle/_ref/libbess_ref.so -- the REAL reference (Mamba413/bess src/*.cpp,
compiled unmodified by oracle/Makefile).  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).
The .so is prebuilt in the build container and shipped to the GPU box; nothing here reads /root/reference."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "libbess_ref.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available() -> bool:
    return os.path.exists(SO_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO_PATH)
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def cv_fold_ids(n: int, K: int, seed: int = 123) -> np.ndarray:
    os.environ["BESS_CV_SEED"] = str(seed)
    out = np.zeros(n, dtype=np.int32)
    lib().ref_cv_fold_ids(C.c_int(n), C.c_int(K), _i(out))
    return out


def pywrap_bess(x, y, data_type, weight, is_normal, algorithm_type, model_type, max_iter, exchange_num, path_type,
                is_warm_start, ic_type, is_cv, K, sequence, s_min, s_max, is_screening, screening_size,
                always_select=(), cv_seed=123, lambda_seq=(0.0,)):
    """Calls the reference's pywrap_bess (bess.cpp:218).  Returns dict(beta, coef0, train_loss, ic)."""
    os.environ["BESS_CV_SEED"] = str(cv_seed)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    n, p = x.shape
    g = np.arange(p, dtype=np.int32)
    state = np.zeros(1)
    seq = np.ascontiguousarray(sequence, dtype=np.int32)
    lam = np.ascontiguousarray(lambda_seq, dtype=np.float64)
    alw = np.ascontiguousarray(always_select, dtype=np.int32)
    beta = np.zeros(p)
    s1 = [np.zeros(1) for _ in range(7)]
    A_out = np.zeros(p, dtype=np.int32)
    l_out = np.zeros(1, dtype=np.int32)
    b = C.c_bool
    i = C.c_int
    dbl = C.c_double
    lib().ref_pywrap_bess(
        _d(x), i(n), i(p), _d(y), i(n), i(data_type), _d(w), i(n), b(is_normal), i(algorithm_type), i(model_type),
        i(max_iter), i(exchange_num), i(path_type), b(is_warm_start), i(ic_type), b(is_cv), i(K), _i(g), i(p),
        _d(state), i(1), _i(seq), i(len(seq)), _d(lam), i(len(lam)), i(s_min), i(s_max), i(10), dbl(10.0), dbl(0.0), dbl(0.0),
        i(len(lam)), b(is_screening), i(screening_size), i(1), _i(alw), i(len(alw)), dbl(1.1),
        _d(beta), i(p), _d(s1[0]), i(1), _d(s1[1]), i(1), _d(s1[2]), i(1), _d(s1[3]), _d(s1[4]), i(1), _d(s1[5]), i(1),
        _d(s1[6]), i(1), _i(A_out), i(p), _i(l_out))
    return dict(beta=beta, coef0=float(s1[0][0]), train_loss=float(s1[1][0]), ic=float(s1[2][0]))


def max_k(vec, k):
    """The reference's max_k (utilities.cpp:179-188) on `vec`: k largest, ascending, boundary ties as std::nth_element
    leaves them."""
    v = np.ascontiguousarray(vec, dtype=np.float64)
    out = np.zeros(k, dtype=np.int32)
    lib().ref_max_k(_d(v), C.c_int(v.size), C.c_int(int(k)), _i(out))
    return out


def screening(x, y, weight, model_type, screening_size):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    n, p = x.shape
    out = np.zeros(screening_size, dtype=np.int32)
    lib().ref_screening(_d(x), C.c_int(n), C.c_int(p), _d(y), _d(w), C.c_int(model_type), C.c_int(screening_size), _i(out))
    return out


def single_fit(x, y, weight, data_type, is_normal, model_type, max_iter, T0, train_mask, beta_init, coef0_init):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    n, p = x.shape
    tm = np.ascontiguousarray(train_mask, dtype=np.int32)
    bi = np.ascontiguousarray(beta_init, dtype=np.float64)
    beta = np.zeros(p)
    c0 = C.c_double(0.0)
    l = C.c_int(0)
    lib().ref_single_fit(_d(x), C.c_int(n), C.c_int(p), _d(y), _d(w), C.c_int(data_type), C.c_bool(is_normal),
                         C.c_int(model_type), C.c_int(max_iter), C.c_int(T0), _i(tm), C.c_int(len(tm)), _d(bi),
                         C.c_double(coef0_init), _d(beta), C.byref(c0), C.byref(l))
    return dict(beta=beta, coef0=c0.value, l=l.value)


def seq_trace(x, y, weight, data_type, is_normal, model_type, max_iter, is_warm_start, ic_type, is_cv, K, sequence,
              cv_seed=123):
    os.environ["BESS_CV_SEED"] = str(cv_seed)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    n, p = x.shape
    seq = np.ascontiguousarray(sequence, dtype=np.int32)
    L = len(seq)
    beta_all = np.zeros((L, p))
    coef0_all, loss_all, ic_all = np.zeros(L), np.zeros(L), np.zeros(L)
    l_all = np.zeros(L, dtype=np.int32)
    lib().ref_seq_trace(_d(x), C.c_int(n), C.c_int(p), _d(y), _d(w), C.c_int(data_type), C.c_bool(is_normal),
                        C.c_int(model_type), C.c_int(max_iter), C.c_bool(is_warm_start), C.c_int(ic_type),
                        C.c_bool(is_cv), C.c_int(K), _i(seq), C.c_int(L), _d(beta_all), _d(coef0_all), _d(loss_all),
                        _d(ic_all), _i(l_all))
    return dict(beta_all=beta_all, coef0_all=coef0_all, loss_all=loss_all, ic_all=ic_all, l_all=l_all)


def bess_lambda(x, y, data_type, weight, is_normal, algorithm_type, model_type, max_iter, path_type, is_warm_start, ic_type,
                is_cv, K, sequence, s_min, s_max, lambda_seq=(0.0,), lambda_min=0.0, lambda_max=0.0, n_lambda=100,
                powell_path=1, is_screening=False, screening_size=1, always_select=(), cv_seed=123, g_index=None):
    """bessCpp (bess.cpp:37) incl. the chosen ridge level: sequential lambda grids and the pgs_path Powell search
    (path.cpp:1138-1309; taken when path_type == 2 and algorithm_type is 5 or 3, bess.cpp:169-176)."""
    os.environ["BESS_CV_SEED"] = str(cv_seed)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    n, p = x.shape
    seq = np.ascontiguousarray(sequence, dtype=np.int32)
    lam = np.ascontiguousarray(lambda_seq, dtype=np.float64)
    alw = np.ascontiguousarray(always_select, dtype=np.int32)
    gi = np.ascontiguousarray(np.arange(p) if g_index is None else g_index, dtype=np.int32)
    beta = np.zeros(p)
    c0, tl, ic, lo = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
    i, b, dbl = C.c_int, C.c_bool, C.c_double
    lib().ref_bess_lambda(_d(x), i(n), i(p), _d(y), i(data_type), _d(w), b(is_normal), i(algorithm_type), i(model_type),
                          i(max_iter), i(path_type), b(is_warm_start), i(ic_type), b(is_cv), i(K), _i(seq), i(len(seq)),
                          _d(lam), i(len(lam)), i(s_min), i(s_max), dbl(lambda_min), dbl(lambda_max), i(n_lambda),
                          b(is_screening), i(screening_size), i(powell_path), _i(alw), i(len(alw)), _i(gi), i(len(gi)),
                          _d(beta), C.byref(c0), C.byref(tl), C.byref(ic), C.byref(lo))
    return dict(beta=beta, coef0=c0.value, train_loss=tl.value, ic=ic.value, lambda_=lo.value)


def pgs_trace(x, y, data_type, weight, is_normal, algorithm_type, model_type, max_iter, is_warm_start, ic_type, is_cv, K,
              s_min, s_max, lambda_min, lambda_max, n_lambda=100, powell_path=1, cv_seed=123, max_rec=20000):
    """pgs_path (path.cpp:1138-1309) with every Algorithm::fit logged: ``fits`` rows are
    (sparsity level, lambda, #train rows, coef0_init, nnz(beta_init)) in call order (full fits and CV fold fits)."""
    os.environ["BESS_CV_SEED"] = str(cv_seed)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    n, p = x.shape
    beta = np.zeros(p)
    rec = np.zeros((max_rec, 5))
    c0, tl, ic, lo = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
    i, b, dbl = C.c_int, C.c_bool, C.c_double
    nf = lib().ref_pgs_trace(_d(x), i(n), i(p), _d(y), _d(w), i(data_type), b(is_normal), i(algorithm_type), i(model_type),
                             i(max_iter), b(is_warm_start), i(ic_type), b(is_cv), i(K), i(s_min), i(s_max), dbl(lambda_min),
                             dbl(lambda_max), i(n_lambda), i(powell_path), _d(beta), C.byref(c0), C.byref(tl), C.byref(ic),
                             C.byref(lo), _d(rec), i(max_rec))
    return dict(beta=beta, coef0=c0.value, train_loss=tl.value, ic=ic.value, lambda_=lo.value, fits=rec[:min(nf, max_rec)],
                n_fits=nf)
