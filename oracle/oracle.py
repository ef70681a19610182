"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/pqa_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The product package (probqa_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpqa_oracle.so")


class RatedTarget(C.Structure):
    _fields_ = [("iTarget", C.c_int64), ("prob", C.c_double)]


class AnsweredQuestion(C.Structure):
    _fields_ = [("iQuestion", C.c_int64), ("iAnswer", C.c_int64)]


class V4(C.Structure):
    _fields_ = [("sum", C.c_double * 4), ("corr", C.c_double * 4)]


def build(force=False):
    src = os.path.join(_HERE, "pqa_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpqa_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ora_log2hot.restype = C.c_double
        L.ora_log2hot.argtypes = [C.c_double]
        L.ora_log2hot_table.restype = _dp
        L.ora_v4_precise_sum.restype = C.c_double
        L.ora_v4_precise_sum.argtypes = [C.POINTER(V4)]
        L.ora_v4_full_sum.restype = C.c_double
        L.ora_v4_full_sum.argtypes = [C.POINTER(V4)]
        L.ora_v4_pair_sum.restype = C.c_double
        L.ora_v4_pair_sum.argtypes = [C.POINTER(V4), C.POINTER(V4), _dp]
        L.ora_v4_add.argtypes = [C.POINTER(V4), _dp]
        L.ora_v4_add_at.argtypes = [C.POINTER(V4), C.c_int, C.c_double]
        L.ora_calc_split.restype = C.c_int64
        L.ora_calc_split.argtypes = [C.c_int64, C.c_int64, _i64p]
        L.ora_start_quiz.argtypes = [_dp, _u8p, C.c_int64, C.c_int64, _dp]
        L.ora_record_answer.argtypes = [_dp, _dp, _u8p, C.c_int64, C.c_int64, _dp]
        L.ora_resume_quiz.restype = C.c_int
        L.ora_resume_quiz.argtypes = [_dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _u8p,
                                      C.POINTER(AnsweredQuestion), C.c_int64, _dp]
        L.ora_eval_question.restype = C.c_double
        L.ora_eval_question.argtypes = [_dp, _dp, C.c_int64, _dp, _u8p, C.c_int64, C.c_int64, C.c_int64,
                                        _dp, _dp, _dp, _dp, _dp]
        L.ora_eval_questions.restype = C.c_int64
        L.ora_eval_questions.argtypes = [_dp, _dp, C.c_int64, _dp, _u8p, _u8p, _u8p, C.c_int64, C.c_int64,
                                         C.c_int64, C.c_int64, C.c_int, _dp, _dp, _dp, _i64p]
        L.ora_select_question.restype = C.c_int64
        L.ora_select_question.argtypes = [_dp, _dp, _i64p, C.c_int64, C.c_int64, C.c_uint64, _u8p, _u8p]
        L.ora_find_nearest_question.restype = C.c_int64
        L.ora_find_nearest_question.argtypes = [C.c_int64, C.c_int64, _u8p, _u8p]
        L.ora_list_top_targets.restype = C.c_int64
        L.ora_list_top_targets.argtypes = [_dp, _u8p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(RatedTarget)]
        L.ora_would_use_radix.restype = C.c_int
        L.ora_would_use_radix.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.ora_record_quiz_target.argtypes = [_dp, _dp, _dp, C.c_int64, C.c_int64, C.POINTER(AnsweredQuestion),
                                             C.c_int64, C.c_int64, C.c_double]
        L.ora_train.argtypes = [_dp, _dp, _dp, C.c_int64, C.c_int64, C.POINTER(AnsweredQuestion),
                                C.c_int64, C.c_int64, C.c_double, C.c_int64]
        L.ora_make_heap.argtypes = [C.POINTER(RatedTarget), C.c_int64]
        L.ora_pop_heap.argtypes = [C.POINTER(RatedTarget), C.c_int64]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _bits(a):
    return None if a is None else a.ctypes.data_as(_u8p)


def pack_bits(flags):
    """bool array -> byte bitmap (bit x at byte x>>3, bit x&7), or None if flags is None."""
    if flags is None:
        return None
    return np.packbits(np.asarray(flags, dtype=np.uint8), bitorder="little")


def log2hot(x):
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    L = lib()
    return np.array([L.ora_log2hot(float(v)) for v in x])


def log2hot_table():
    p = lib().ora_log2hot_table()
    return np.ctypeslib.as_array(p, shape=(1024,)).copy()


def calc_split(n, w):
    b = np.zeros(max(int(w), 1), dtype=np.int64)
    k = lib().ora_calc_split(n, w, b.ctypes.data_as(_i64p))
    return b[:k].copy()


def start_quiz(vB, W, tgaps=None):
    vB = np.ascontiguousarray(vB, dtype=np.float64)
    prior = np.empty_like(vB)
    g = pack_bits(tgaps)
    lib().ora_start_quiz(_d(vB), _bits(g), vB.size, W, _d(prior))
    return prior


def record_answer(prior, sArow, mDrow, W, tgaps=None):
    """W is the loose worker count max(1, hwc-1). Returns the new prior."""
    prior = np.array(prior, dtype=np.float64, copy=True)
    sArow = np.ascontiguousarray(sArow, dtype=np.float64)
    mDrow = np.ascontiguousarray(mDrow, dtype=np.float64)
    g = pack_bits(tgaps)
    lib().ora_record_answer(_d(sArow), _d(mDrow), _bits(g), prior.size, W, _d(prior))
    return prior


def resume_quiz(sA, mD, vB, aqs, W, tgaps=None):
    """ResumeQuiz priors for the answered questions aqs = [(q, a), ...] (len >= 1). Raises on the I64Underflow case."""
    sA = np.ascontiguousarray(sA, dtype=np.float64); mD = np.ascontiguousarray(mD, dtype=np.float64)
    vB = np.ascontiguousarray(vB, dtype=np.float64)
    Q, K, T = sA.shape
    prior = np.empty(T)
    g = pack_bits(tgaps)
    rc = lib().ora_resume_quiz(_d(sA), _d(mD), _d(vB), T, K, T, W, _bits(g), _aq_array(aqs), len(aqs), _d(prior))
    if rc != 0:
        raise OverflowError("I64Underflow (CpuEngine.cpp:316-319)")
    return prior


def eval_question(sAi, mDi, prior, nValidTargets=None, tgaps=None):
    """sAi: [K, T] array, mDi: [T]. Returns dict(priority, W, H, V, lack, totW)."""
    sAi = np.ascontiguousarray(sAi, dtype=np.float64)
    mDi = np.ascontiguousarray(mDi, dtype=np.float64)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    K, T = sAi.shape
    g = pack_bits(tgaps)
    if nValidTargets is None:
        nValidTargets = T - (0 if tgaps is None else int(np.sum(tgaps)))
    Wk = np.empty(K); Hk = np.empty(K); Vk = np.empty(K)
    lack = C.c_double(); totW = C.c_double()
    pr = lib().ora_eval_question(_d(sAi), _d(mDi), T, _d(prior), _bits(g), K, T, nValidTargets,
                                 _d(Wk), _d(Hk), _d(Vk), C.byref(lack), C.byref(totW))
    return dict(priority=pr, W=Wk, H=Hk, V=Vk, lack=lack.value, totW=totW.value)


def eval_questions(sA, mD, prior, W, asked=None, qgaps=None, tgaps=None, nThreads=1):
    """sA: [Q, K, T], mD: [Q, T]. Returns dict(runLength[Q], priority[Q], grand[nChunks], bounds[nChunks])."""
    sA = np.ascontiguousarray(sA, dtype=np.float64)
    mD = np.ascontiguousarray(mD, dtype=np.float64)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    Q, K, T = sA.shape
    a, qg, tg = pack_bits(asked), pack_bits(qgaps), pack_bits(tgaps)
    run = np.empty(Q); pri = np.empty(Q)
    grand = np.empty(8 * W); bounds = np.zeros(8 * W, dtype=np.int64)
    n = lib().ora_eval_questions(_d(sA), _d(mD), T, _d(prior), _bits(a), _bits(qg), _bits(tg), Q, K, T, W,
                                 nThreads, _d(run), _d(pri), _d(grand), bounds.ctypes.data_as(_i64p))
    return dict(runLength=run, priority=pri, grand=grand[:n].copy(), bounds=bounds[:n].copy())


def select_question(ev, Q, rnd, asked=None, qgaps=None):
    a, qg = pack_bits(asked), pack_bits(qgaps)
    run = np.ascontiguousarray(ev["runLength"]); grand = np.ascontiguousarray(ev["grand"])
    bounds = np.ascontiguousarray(ev["bounds"], dtype=np.int64)
    return lib().ora_select_question(_d(run), _d(grand), bounds.ctypes.data_as(_i64p), bounds.size, Q,
                                     C.c_uint64(int(rnd)), _bits(a), _bits(qg))


def find_nearest_question(iMiddle, Q, asked=None, qgaps=None):
    a, qg = pack_bits(asked), pack_bits(qgaps)
    return lib().ora_find_nearest_question(iMiddle, Q, _bits(a), _bits(qg))


def list_top_targets(prior, W, maxCount, tgaps=None):
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    g = pack_bits(tgaps)
    dest = (RatedTarget * max(int(maxCount), 1))()
    n = lib().ora_list_top_targets(_d(prior), _bits(g), prior.size, W, maxCount, dest)
    return [(dest[i].iTarget, dest[i].prob) for i in range(n)]


def _aq_array(aqs):
    arr = (AnsweredQuestion * max(len(aqs), 1))()
    for i, (q, a) in enumerate(aqs):
        arr[i].iQuestion = int(q); arr[i].iAnswer = int(a)
    return arr


def record_quiz_target(sA, mD, vB, aqs, iTarget, amount=1.0):
    """In-place update of C-contiguous sA[Q,K,T], mD[Q,T], vB[T]."""
    Q, K, T = sA.shape
    assert sA.flags.c_contiguous and mD.flags.c_contiguous and vB.flags.c_contiguous
    lib().ora_record_quiz_target(_d(sA), _d(mD), _d(vB), T, K, _aq_array(aqs), len(aqs), iTarget, amount)


def train(sA, mD, vB, aqs, iTarget, amount=1.0, W=1):
    Q, K, T = sA.shape
    assert sA.flags.c_contiguous and mD.flags.c_contiguous and vB.flags.c_contiguous
    lib().ora_train(_d(sA), _d(mD), _d(vB), T, K, _aq_array(aqs), len(aqs), iTarget, amount, W)
