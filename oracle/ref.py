"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libpqa_ref.so -- the reference's OWN hot-path code
(/root/reference/ProbQA, compiled by oracle/build_ref.sh). Used by tests/ to pin the oracle restatement and by
bench.py's reference arm. Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpqa_ref.so")

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)


class RatedTarget(C.Structure):
    _fields_ = [("iTarget", C.c_int64), ("prob", C.c_double)]


class HeadItem(C.Structure):
    _fields_ = [("prob", C.c_double), ("iSource", C.c_int64)]


class AnsweredQuestion(C.Structure):
    _fields_ = [("iQuestion", C.c_int64), ("iAnswer", C.c_int64)]


def build():
    """(Re)build from /root/reference when it is present; on the GPU box the prebuilt .so is used."""
    subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return LIB_PATH


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libpqa_ref.so missing: run oracle/build_ref.sh where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.ref_log2hot.argtypes = [_dp, _dp, C.c_int64]
        L.ref_v4_accumulate.argtypes = [_dp, C.c_int64, _dp, _dp]
        L.ref_v4_pair.argtypes = [_dp, _dp, C.c_int64, _dp, _dp]
        L.ref_v4_pair_at.argtypes = [_dp, _dp, C.c_int64, _dp, _dp]
        L.ref_kahan_scalar.restype = C.c_double
        L.ref_kahan_scalar.argtypes = [_dp, C.c_int64]
        for f in ("ref_make_heap", "ref_pop_heap"):
            getattr(L, f).argtypes = [C.POINTER(RatedTarget), C.c_int64]
        for f in ("ref_head_make_heap", "ref_head_pop_heap", "ref_head_down"):
            getattr(L, f).argtypes = [C.POINTER(HeadItem), C.c_int64]
        L.ref_calc_split.restype = C.c_int64
        L.ref_calc_split.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_size_t)]
        L.ref_engine_create.restype = C.c_void_p
        L.ref_engine_create.argtypes = [C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _dp, _u8p, _u8p, C.c_int64]
        L.ref_engine_destroy.argtypes = [C.c_void_p]
        L.ref_engine_set_os_threads.argtypes = [C.c_void_p, C.c_int64]
        L.ref_log_count.restype = C.c_longlong
        L.ref_engine_read_kb.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ref_start_quiz.argtypes = [C.c_void_p, _dp]
        L.ref_record_answer.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64]
        L.ref_eval_questions.restype = C.c_int64
        L.ref_eval_questions.argtypes = [C.c_void_p, _dp, _u8p, _dp, _dp, _i64p]
        L.ref_list_top_targets.restype = C.c_int64
        L.ref_resume_quiz.restype = C.c_int64
        L.ref_resume_quiz.argtypes = [C.c_void_p, C.POINTER(AnsweredQuestion), C.c_int64, _dp]
        L.ref_list_top_targets.argtypes = [C.c_void_p, _dp, C.c_int64, C.POINTER(RatedTarget)]
        L.ref_record_quiz_target.argtypes = [C.c_void_p, C.POINTER(AnsweredQuestion), C.c_int64, C.c_int64, C.c_double]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _bits(flags):
    if flags is None:
        return None, None
    b = np.packbits(np.asarray(flags, dtype=np.uint8), bitorder="little")
    b = np.concatenate([b, np.zeros(8, dtype=np.uint8)])
    return b, b.ctypes.data_as(_u8p)


def log2hot(x):
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    n = x.size
    pad = (-n) % 4
    xp = np.concatenate([x, np.ones(pad)])
    out = np.empty_like(xp)
    lib().ref_log2hot(_d(xp), _d(out), xp.size)
    return out[:n]


def v4_accumulate(values):
    """values: [nVects, 4]. Returns (PreciseSum, GetFullSum)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    ps, fs = C.c_double(), C.c_double()
    lib().ref_v4_accumulate(_d(v), v.shape[0], C.byref(ps), C.byref(fs))
    return ps.value, fs.value


def v4_pair(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    sa, sb = C.c_double(), C.c_double()
    lib().ref_v4_pair(_d(a), _d(b), a.shape[0], C.byref(sa), C.byref(sb))
    return sa.value, sb.value


def v4_pair_at(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel(); b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    sa, sb = C.c_double(), C.c_double()
    lib().ref_v4_pair_at(_d(a), _d(b), a.size, C.byref(sa), C.byref(sb))
    return sa.value, sb.value


def kahan_scalar(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().ref_kahan_scalar(_d(v), v.size)


def calc_split(n, w):
    b = (C.c_size_t * max(int(w), 1))()
    k = lib().ref_calc_split(n, w, b)
    return np.array(b[:k], dtype=np.int64)


class RefEngine:
    """The reference's subtask bodies over a stub engine shell. nWorkers plays hardware_concurrency()."""

    def __init__(self, sA, mD, vB, nWorkers, qgaps=None, tgaps=None, osThreads=1):
        sA = np.ascontiguousarray(sA, dtype=np.float64); mD = np.ascontiguousarray(mD, dtype=np.float64)
        vB = np.ascontiguousarray(vB, dtype=np.float64)
        self.Q, self.K, self.T = sA.shape
        self.W = nWorkers
        self._qg, qgp = _bits(qgaps)
        self._tg, tgp = _bits(tgaps)
        self.h = C.c_void_p(lib().ref_engine_create(self.Q, self.K, self.T, _d(sA), _d(mD), _d(vB), qgp, tgp, nWorkers))
        if osThreads > 1:
            lib().ref_engine_set_os_threads(self.h, osThreads)

    def close(self):
        if self.h:
            lib().ref_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def read_kb(self):
        sA = np.empty((self.Q, self.K, self.T)); mD = np.empty((self.Q, self.T)); vB = np.empty(self.T)
        lib().ref_engine_read_kb(self.h, _d(sA), _d(mD), _d(vB))
        return sA, mD, vB

    def start_quiz(self):
        p = np.empty(self.T)
        lib().ref_start_quiz(self.h, _d(p))
        return p

    def record_answer(self, prior, q, a):
        p = np.array(prior, dtype=np.float64, copy=True)
        lib().ref_record_answer(self.h, _d(p), q, a)
        return p

    def resume_quiz(self, aqs):
        arr = (AnsweredQuestion * max(len(aqs), 1))()
        for i, (q, a) in enumerate(aqs):
            arr[i].iQuestion = int(q); arr[i].iAnswer = int(a)
        p = np.empty(self.T)
        rc = lib().ref_resume_quiz(self.h, arr, len(aqs), _d(p))
        if rc != 0:
            raise OverflowError("I64Underflow")
        return p

    def eval_questions(self, prior, asked=None):
        prior = np.ascontiguousarray(prior, dtype=np.float64)
        _a, ap = _bits(asked)
        run = np.empty(self.Q); grand = np.empty(8 * self.W); bounds = np.zeros(8 * self.W, dtype=np.int64)
        n = lib().ref_eval_questions(self.h, _d(prior), ap, _d(run), _d(grand), bounds.ctypes.data_as(_i64p))
        return dict(runLength=run, grand=grand[:n].copy(), bounds=bounds[:n].copy())

    def list_top_targets(self, prior, maxCount):
        prior = np.ascontiguousarray(prior, dtype=np.float64)
        dest = (RatedTarget * max(int(maxCount), 1))()
        n = lib().ref_list_top_targets(self.h, _d(prior), maxCount, dest)
        return [(dest[i].iTarget, dest[i].prob) for i in range(n)]

    def record_quiz_target(self, aqs, iTarget, amount=1.0):
        arr = (AnsweredQuestion * max(len(aqs), 1))()
        for i, (q, a) in enumerate(aqs):
            arr[i].iQuestion = int(q); arr[i].iAnswer = int(a)
        lib().ref_record_quiz_target(self.h, arr, len(aqs), iTarget, amount)
