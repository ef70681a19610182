"""ctypes host binding of libPqaCore.so (the B200 engine), mirroring the class and method names of the reference's
own Python client (Interop/Python/ProbQAInterop/ProbQA.py:300-760: PqaEngineFactory.create_cpu_engine, PqaEngine.
start_quiz / next_question / record_answer / list_top_targets / record_quiz_target / train / release_quiz ...) so that
code written against the reference binding runs unchanged, plus numpy-array batch methods over the additive entry
points of include/PqaB200Ext.h.

There is no fallback: if the shared library is missing or no CUDA device is visible, construction raises.
Nothing in here imports or calls oracle/.
"""
import ctypes as C
import os
from enum import Enum
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libPqaCore.so")


class CiEngineDefinition(C.Structure):  # PqaCInterop.h:10-19, pack 8
    _pack_ = 8
    _fields_ = [("nAnswers", C.c_int64), ("nQuestions", C.c_int64), ("nTargets", C.c_int64),
                ("precType", C.c_uint8), ("precExponent", C.c_uint16), ("precMantissa", C.c_uint32),
                ("initAmount", C.c_double), ("memPoolMaxBytes", C.c_uint64)]


class CiAnsweredQuestion(C.Structure):  # :21-24
    _pack_ = 8
    _fields_ = [("iQuestion", C.c_int64), ("iAnswer", C.c_int64)]


class CiEngineDimensions(C.Structure):  # :26-30
    _pack_ = 8
    _fields_ = [("nAnswers", C.c_int64), ("nQuestions", C.c_int64), ("nTargets", C.c_int64)]


class CiRatedTarget(C.Structure):  # :32-35
    _pack_ = 8
    _fields_ = [("iTarget", C.c_int64), ("prob", C.c_double)]


class CiAddQorTParam(C.Structure):  # :37-40
    _pack_ = 8
    _fields_ = [("index", C.c_int64), ("initAmount", C.c_double)]


class CiB200GroupOptions(C.Structure):  # PqaB200Ext.h
    _pack_ = 8
    _fields_ = [("axis", C.c_int32), ("nShards", C.c_int32), ("devices", C.c_int32 * 8), ("exactOrder", C.c_int32),
                ("maxBatch", C.c_int64)]


class CiB200Options(C.Structure):  # PqaB200Ext.h
    _pack_ = 8
    _fields_ = [("device", C.c_int32), ("emulatedWorkers", C.c_int32), ("rngSeed", C.c_uint64),
                ("initialQuizCapacity", C.c_int64), ("questionShardFirst", C.c_int64), ("questionShardCount", C.c_int64),
                ("targetShardFirst", C.c_int64), ("targetShardCount", C.c_int64)]


RATED_DTYPE = np.dtype([("iTarget", np.int64), ("prob", np.float64)])

_vp, _i64, _u64, _dbl = C.c_void_p, C.c_int64, C.c_uint64, C.c_double
_pi64, _pu64, _pd = C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol declared in include/PqaCInterop.h and include/PqaB200Ext.h
SIGNATURES = {
    "CiDebugBreak": (None, []),
    "Logger_Init": (C.c_uint8, [_pvp, C.c_char_p]),
    "CiReleaseString": (None, [_vp]),
    "CiGetPqaEngineFactory": (_vp, []),
    "PqaEngineFactory_CreateCpuEngine": (_vp, [_vp, _pvp, C.POINTER(CiEngineDefinition)]),
    "PqaEngineFactory_LoadCpuEngine": (_vp, [_vp, _pvp, C.c_char_p, _u64]),
    "CiReleasePqaError": (None, [_vp]),
    "PqaError_ToString": (_vp, [_vp, C.c_uint8]),
    "CiReleasePqaEngine": (None, [_vp]),
    "PqaEngine_Train": (_vp, [_vp, _i64, C.POINTER(CiAnsweredQuestion), _i64, _dbl]),
    "PqaEngine_QuestionPermFromComp": (C.c_uint8, [_vp, _i64, _pi64]),
    "PqaEngine_QuestionCompFromPerm": (C.c_uint8, [_vp, _i64, _pi64]),
    "PqaEngine_TargetPermFromComp": (C.c_uint8, [_vp, _i64, _pi64]),
    "PqaEngine_TargetCompFromPerm": (C.c_uint8, [_vp, _i64, _pi64]),
    "PqaEngine_QuizPermFromComp": (C.c_uint8, [_vp, _i64, _pi64]),
    "PqaEngine_QuizCompFromPerm": (C.c_uint8, [_vp, _i64, _pi64]),
    "PqaEngine_EnsurePermQuizGreater": (C.c_uint8, [_vp, _i64]),
    "PqaEngine_RemapQuizPermId": (C.c_uint8, [_vp, _i64, _i64]),
    "PqaEngine_GetTotalQuestionsAsked": (_u64, [_vp, _pvp]),
    "PqaEngine_CopyDims": (C.c_uint8, [_vp, C.POINTER(CiEngineDimensions)]),
    "PqaEngine_StartQuiz": (_i64, [_vp, _pvp]),
    "PqaEngine_ResumeQuiz": (_i64, [_vp, _pvp, _i64, C.POINTER(CiAnsweredQuestion)]),
    "PqaEngine_NextQuestion": (_i64, [_vp, _pvp, _i64]),
    "PqaEngine_RecordAnswer": (_vp, [_vp, _i64, _i64]),
    "PqaEngine_ClearOldQuizzes": (_vp, [_vp, _i64, _dbl]),
    "PqaEngine_GetActiveQuestionId": (_i64, [_vp, _pvp, _i64]),
    "PqaEngine_SetActiveQuestion": (_vp, [_vp, _i64, _i64]),
    "PqaEngine_ListTopTargets": (_i64, [_vp, _pvp, _i64, _i64, C.POINTER(CiRatedTarget)]),
    "PqaEngine_RecordQuizTarget": (_vp, [_vp, _i64, _i64, _dbl]),
    "PqaEngine_ReleaseQuiz": (_vp, [_vp, _i64]),
    "PqaEngine_SaveKB": (_vp, [_vp, C.c_char_p, C.c_uint8]),
    "PqaEngine_StartMaintenance": (_vp, [_vp, C.c_bool]),
    "PqaEngine_FinishMaintenance": (_vp, [_vp]),
    "PqaEngine_AddQsTs": (_vp, [_vp, _i64, _vp, _i64, _vp]),
    "PqaEngine_RemoveQuestions": (_vp, [_vp, _i64, _pi64]),
    "PqaEngine_RemoveTargets": (_vp, [_vp, _i64, _pi64]),
    "PqaEngine_Compact": (_vp, [_vp, _pi64, C.POINTER(_pi64), _pi64, C.POINTER(_pi64)]),
    "CiReleaseCompaction": (None, [_pi64]),
    "PqaEngine_Shutdown": (_vp, [_vp, C.c_char_p]),
    "PqaEngine_SetLogger": (_vp, [_vp, _vp]),
    # ---- PqaB200Ext.h
    "PqaB200_CreateEngine": (_vp, [_pvp, C.POINTER(CiEngineDefinition), C.POINTER(CiB200Options)]),
    "PqaB200_LoadEngine": (_vp, [_pvp, C.c_char_p, C.POINTER(CiB200Options)]),
    "PqaB200_CreateShardedEngine": (_vp, [_pvp, C.POINTER(CiEngineDefinition), C.POINTER(CiB200Options), C.POINTER(CiB200GroupOptions)]),
    "PqaB200_LoadShardedEngine": (_vp, [_pvp, C.c_char_p, C.POINTER(CiB200Options), C.POINTER(CiB200GroupOptions)]),
    "PqaB200_GetShardCount": (C.c_int32, [_vp]),
    "PqaB200_HostLogicSelfTest": (_vp, []),
    "PqaB200_SaveKBShard": (_vp, [_vp, C.c_char_p, C.c_int32]),
    "PqaB200_GetEmulatedWorkers": (C.c_int32, [_vp]),
    "PqaB200_GetDevice": (C.c_int32, [_vp]),
    "PqaB200_BuildInfo": (C.c_char_p, []),
    "PqaEngine_CopyATargets": (_vp, [_vp, _i64, _i64, _i64, _pd]),
    "PqaEngine_CopyDTargets": (_vp, [_vp, _i64, _i64, _pd]),
    "PqaEngine_CopyBTargets": (_vp, [_vp, _i64, _pd]),
    "PqaB200_UploadKB": (_vp, [_vp, _pd, _pd, _pd]),
    "PqaB200_DownloadKB": (_vp, [_vp, _pd, _pd, _pd]),
    "PqaEngine_StartQuizBatch": (_vp, [_vp, _i64, _pi64]),
    "PqaEngine_ResumeQuizBatch": (_vp, [_vp, _i64, _pi64, C.POINTER(CiAnsweredQuestion), _pi64]),
    "PqaEngine_NextQuestionBatch": (_vp, [_vp, _i64, _pi64, _pu64, _pi64, _pvp]),
    "PqaEngine_RecordAnswerBatch": (_vp, [_vp, _i64, _pi64, _pi64]),
    "PqaEngine_SetActiveQuestionBatch": (_vp, [_vp, _i64, _pi64, _pi64]),
    "PqaEngine_ListTopTargetsBatch": (_vp, [_vp, _i64, _pi64, _i64, _vp, _pi64]),
    "PqaEngine_RecordQuizTargetBatch": (_vp, [_vp, _i64, _pi64, _pi64, _pd]),
    "PqaEngine_ReleaseQuizBatch": (_vp, [_vp, _i64, _pi64]),
    "PqaB200_CopyQuizPriors": (_vp, [_vp, _i64, _pd]),
    "PqaB200_SetQuizPriors": (_vp, [_vp, _i64, _pd]),
    "PqaB200_EvalQuestions": (_vp, [_vp, _i64, _pi64, _pd, _pd, _pd, _pi64]),
    "PqaB200_EvalQuestionsDetailed": (_vp, [_vp, _i64, _pd, _pd, _pd, _pd, _pd]),
    "PqaB200_P2PLastPhaseMs": (_vp, [_vp, _pd]),
    "PqaB200_AnomalyCounts": (_vp, [_vp, _pu64]),
    "PqaB200_EvalQuestionsDetailedBatch": (_vp, [_vp, _i64, _pi64, _pd, _pd, _pd, _pd, _pd]),
    "PqaB200_SetEvalKernel": (_vp, [_vp, C.c_int32]),
    "PqaB200_SetEvalTuning": (_vp, [_vp, C.c_int32, _i64, _i64, C.c_int32]),
    "PqaB200_ShardEval": (_vp, [_vp, _i64, _pi64]),
    "PqaB200_ShardSelect": (_vp, [_vp, _i64, _pi64, _pu64, _pi64, _pvp]),
    "PqaB200_ShardRecordAnswerBegin": (_vp, [_vp, _i64, _pi64, _pi64]),
    "PqaB200_ShardRecordAnswerEnd": (_vp, [_vp, _i64, _pi64]),
    "PqaB200_ShardBuffer": (_vp, [_vp, C.c_int32, _pvp, _pi64]),
    "PqaB200_GetQuestionShard": (_vp, [_vp, _pi64, _pi64]),
    "PqaB200_TShardEvalW": (_vp, [_vp, _i64, _pi64]),
    "PqaB200_TShardEvalHVL": (_vp, [_vp, _i64, _pi64]),
    "PqaB200_TShardPriority": (_vp, [_vp, _i64, _pi64]),
    "PqaB200_GetTargetShard": (_vp, [_vp, _pi64, _pi64]),
    "PqaB200_FillBinarySearchKB": (_vp, [_vp, C.c_double]),
    "PqaB200_P2PInit": (_vp, [_vp, C.c_int32, C.c_int32, _i64, _pvp, _pi64]),
    "PqaB200_P2PExportHandle": (_vp, [_vp, C.c_char_p]),
    "PqaB200_P2POpenHandle": (_vp, [_vp, C.c_char_p, _pvp]),
    "PqaB200_P2PConnect": (_vp, [_vp, _pvp]),
    "PqaB200_P2PNextQuestionBegin": (_vp, [_vp, _i64, _pi64, _pu64]),
    "PqaB200_P2PNextQuestionEnd": (_vp, [_vp, _i64, _pi64, _pi64, _pvp]),
    "PqaB200_P2PRecordAnswerBegin": (_vp, [_vp, _i64, _pi64, _pi64]),
    "PqaB200_P2PRecordAnswerEnd": (_vp, [_vp]),
    "PqaB200_P2PSetExactOrder": (_vp, [_vp, C.c_int32]),
    "PqaB200_ResidentBind": (_vp, [_vp, _i64, _pi64, _pu64]),
    "PqaB200_ResidentStep": (_vp, [_vp]),
    "PqaB200_ResidentFetch": (_vp, [_vp, _pi64]),
    "PqaB200_ResidentLastEvalMs": (_dbl, [_vp]),
    "PqaB200_Synchronize": (_vp, [_vp]),
    "PqaB200_EventCreate": (_vp, []),
    "PqaB200_EventDestroy": (None, [_vp]),
    "PqaB200_EventRecord": (_vp, [_vp, _vp]),
    "PqaB200_EventSynchronize": (_vp, [_vp]),
    "PqaB200_EventElapsedMs": (_dbl, [_vp, _vp]),
    "PqaB200_KernelLaunchCount": (_u64, [_vp]),
    "PqaB200_FlushL2": (_vp, [_vp]),
}

_lib = None


def load_library(path: str = None):
    """Loads libPqaCore.so and declares every entry point. Raises if the library is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("PQA_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s not found: build it with `python -m probqa_b200.build` (there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class PqaException(Exception):
    def __init__(self, error: "PqaError"):
        super().__init__(error.to_string(True))
        self.error = error


class PrecisionType(Enum):  # Interface/PqaCommon.h:17-24
    NONE = 0
    FLOAT = 1
    FLOAT_PAIR = 2
    DOUBLE = 3
    DOUBLE_PAIR = 4
    ARBITRARY = 5


class AnsweredQuestion:
    def __init__(self, i_question, i_answer):
        self.i_question = i_question
        self.i_answer = i_answer

    def __repr__(self):
        return "[AnsweredQuestion: i_question=%d, i_answer=%d]" % (self.i_question, self.i_answer)


class RatedTarget:
    def __init__(self, i_target: int, prob: float):
        self.i_target = i_target
        self.prob = prob

    def __repr__(self):
        return "[RatedTarget: i_target=%d, prob=%r]" % (self.i_target, self.prob)


class EngineDefinition:
    def __init__(self, n_answers: int, n_questions: int, n_targets: int, init_amount=1.0,
                 prec_type=PrecisionType.DOUBLE, prec_exponent=11, prec_mantissa=53, mem_pool_max_bytes=512 << 20):
        self.n_answers, self.n_questions, self.n_targets = n_answers, n_questions, n_targets
        self.init_amount = init_amount
        self.prec_type, self.prec_exponent, self.prec_mantissa = prec_type, prec_exponent, prec_mantissa
        self.mem_pool_max_bytes = mem_pool_max_bytes

    def to_c(self) -> CiEngineDefinition:
        return CiEngineDefinition(self.n_answers, self.n_questions, self.n_targets, self.prec_type.value,
                                  self.prec_exponent, self.prec_mantissa, self.init_amount, self.mem_pool_max_bytes)


class EngineDimensions:
    def __init__(self, n_answers: int, n_questions: int, n_targets: int):
        self.n_answers, self.n_questions, self.n_targets = n_answers, n_questions, n_targets

    def __repr__(self):
        return "[n_answers=%d, n_questions=%d, n_targets=%d]" % (self.n_answers, self.n_questions, self.n_targets)


class PqaError:
    """Owns a native error object (NULL = success), like ProbQA.py:399-421."""

    @staticmethod
    def factor(c_err):
        return PqaError(c_err) if c_err else None

    def __init__(self, c_err):
        self.c_err = C.c_void_p(c_err) if not isinstance(c_err, C.c_void_p) else c_err

    def __del__(self):
        if getattr(self, "c_err", None) and self.c_err.value and _lib is not None:
            _lib.CiReleasePqaError(self.c_err)
            self.c_err = None

    def __repr__(self):
        return self.to_string(True)

    def to_string(self, with_params: bool) -> str:
        lib = load_library()
        s = lib.PqaError_ToString(self.c_err, 1 if with_params else 0)
        try:
            return C.cast(s, C.c_char_p).value.decode("utf-8", "replace")
        finally:
            lib.CiReleaseString(s)


def _raise_or_return(c_err, throw=True):
    err = PqaError.factor(c_err)
    if err is not None and throw:
        raise PqaException(err)
    return err


def _i64arr(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


class PqaEngine:
    def __init__(self, c_engine):
        self.c_engine = C.c_void_p(c_engine)
        self._lib = load_library()
        d = self.copy_dims()
        self.n_answers, self.n_questions, self.n_targets = d.n_answers, d.n_questions, d.n_targets

    def _refresh_dims(self):
        d = self.copy_dims()
        self.n_answers, self.n_questions, self.n_targets = d.n_answers, d.n_questions, d.n_targets

    def __del__(self):
        self.close()

    def close(self):
        if getattr(self, "c_engine", None) and self.c_engine.value:
            self._lib.CiReleasePqaEngine(self.c_engine)
            self.c_engine = C.c_void_p(None)

    # ---------------------------------------------------------------- reference-binding methods
    @staticmethod
    def to_c_answered_questions(answered_questions: List[AnsweredQuestion]):
        n = len(answered_questions)
        arr = (CiAnsweredQuestion * max(n, 1))()
        for i, aq in enumerate(answered_questions):
            if isinstance(aq, AnsweredQuestion):
                arr[i].iQuestion, arr[i].iAnswer = aq.i_question, aq.i_answer
            else:
                arr[i].iQuestion, arr[i].iAnswer = int(aq[0]), int(aq[1])
        return arr, n

    def train(self, answered_questions, i_target: int, amount: float = 1.0, throw: bool = True):
        arr, n = self.to_c_answered_questions(answered_questions)
        return _raise_or_return(self._lib.PqaEngine_Train(self.c_engine, n, arr, i_target, amount), throw)

    def get_total_questions_asked(self) -> int:
        e = C.c_void_p()
        v = self._lib.PqaEngine_GetTotalQuestionsAsked(self.c_engine, C.byref(e))
        _raise_or_return(e.value)
        return v

    def copy_dims(self) -> EngineDimensions:
        d = CiEngineDimensions()
        if not self._lib.PqaEngine_CopyDims(self.c_engine, C.byref(d)):
            raise RuntimeError("PqaEngine_CopyDims failed")
        return EngineDimensions(d.nAnswers, d.nQuestions, d.nTargets)

    def start_quiz(self) -> int:
        e = C.c_void_p()
        q = self._lib.PqaEngine_StartQuiz(self.c_engine, C.byref(e))
        _raise_or_return(e.value)
        return q

    def resume_quiz(self, answered_questions) -> int:
        arr, n = self.to_c_answered_questions(answered_questions)
        e = C.c_void_p()
        q = self._lib.PqaEngine_ResumeQuiz(self.c_engine, C.byref(e), n, arr)
        _raise_or_return(e.value)
        return q

    def next_question(self, i_quiz: int) -> int:
        e = C.c_void_p()
        q = self._lib.PqaEngine_NextQuestion(self.c_engine, C.byref(e), i_quiz)
        _raise_or_return(e.value)
        return q

    def record_answer(self, i_quiz: int, i_answer: int, throw: bool = True):
        return _raise_or_return(self._lib.PqaEngine_RecordAnswer(self.c_engine, i_quiz, i_answer), throw)

    def get_active_question_id(self, i_quiz: int) -> int:
        e = C.c_void_p()
        q = self._lib.PqaEngine_GetActiveQuestionId(self.c_engine, C.byref(e), i_quiz)
        _raise_or_return(e.value)
        return q

    def set_active_question(self, i_quiz: int, i_question: int, throw: bool = True):
        return _raise_or_return(self._lib.PqaEngine_SetActiveQuestion(self.c_engine, i_quiz, i_question), throw)

    def list_top_targets(self, i_quiz: int, max_count: int) -> List[RatedTarget]:
        dest = (CiRatedTarget * max(max_count, 1))()
        e = C.c_void_p()
        n = self._lib.PqaEngine_ListTopTargets(self.c_engine, C.byref(e), i_quiz, max_count, dest)
        _raise_or_return(e.value)
        return [RatedTarget(dest[i].iTarget, dest[i].prob) for i in range(n)]

    def record_quiz_target(self, i_quiz: int, i_target: int, amount: float = 1.0, throw: bool = True):
        return _raise_or_return(self._lib.PqaEngine_RecordQuizTarget(self.c_engine, i_quiz, i_target, amount), throw)

    def release_quiz(self, i_quiz: int, throw: bool = True):
        return _raise_or_return(self._lib.PqaEngine_ReleaseQuiz(self.c_engine, i_quiz), throw)

    def clear_old_quizzes(self, max_count: int, max_age_sec: float, throw: bool = True):
        """ProbQA.py clear_old_quizzes: releases quizzes idle for more than max_age_sec, then the oldest beyond max_count."""
        return _raise_or_return(self._lib.PqaEngine_ClearOldQuizzes(self.c_engine, max_count, float(max_age_sec)), throw)

    def save_kb(self, file_path: str, b_double_buffer: bool = False, throw: bool = True):
        return _raise_or_return(self._lib.PqaEngine_SaveKB(self.c_engine, file_path.encode(), int(b_double_buffer)), throw)

    # maintenance mode and id maps: same names and shapes as ProbQA.py:634-720 / :560-632
    def start_maintenance(self, force_quizzes: bool, throw: bool = True):
        return _raise_or_return(self._lib.PqaEngine_StartMaintenance(self.c_engine, bool(force_quizzes)), throw)

    def finish_maintenance(self, throw: bool = True):
        err = _raise_or_return(self._lib.PqaEngine_FinishMaintenance(self.c_engine), throw)
        self._refresh_dims()
        return err

    def add_qs_ts(self, question_init_amounts, target_init_amounts, throw: bool = True):
        """Adds len(question_init_amounts) questions and len(target_init_amounts) targets; returns the compact ids they
        received (removed ids are reused first), or the error when throw is False."""
        nq, nt = len(question_init_amounts), len(target_init_amounts)
        cq, ct = (CiAddQorTParam * max(nq, 1))(), (CiAddQorTParam * max(nt, 1))()
        for i, a in enumerate(question_init_amounts):
            cq[i].index, cq[i].initAmount = -1, float(a)
        for i, a in enumerate(target_init_amounts):
            ct[i].index, ct[i].initAmount = -1, float(a)
        err = _raise_or_return(self._lib.PqaEngine_AddQsTs(self.c_engine, nq, C.cast(cq, C.c_void_p), nt, C.cast(ct, C.c_void_p)), throw)
        if err is not None:
            return err
        self._refresh_dims()
        return [cq[i].index for i in range(nq)], [ct[i].index for i in range(nt)]

    def remove_questions(self, question_ids, throw: bool = True):
        ids = _i64arr(question_ids)
        return _raise_or_return(self._lib.PqaEngine_RemoveQuestions(self.c_engine, ids.size, _p(ids, _pi64)), throw)

    def remove_targets(self, target_ids, throw: bool = True):
        ids = _i64arr(target_ids)
        return _raise_or_return(self._lib.PqaEngine_RemoveTargets(self.c_engine, ids.size, _p(ids, _pi64)), throw)

    def compact(self):
        """(old_questions, old_targets): new compact id -> the compact id it had before (ProbQA.py:700-720)."""
        nq, nt = C.c_int64(), C.c_int64()
        pq, pt = _pi64(), _pi64()
        _raise_or_return(self._lib.PqaEngine_Compact(self.c_engine, C.byref(nq), C.byref(pq), C.byref(nt), C.byref(pt)))
        try:
            return [pq[i] for i in range(nq.value)], [pt[i] for i in range(nt.value)]
        finally:
            self._lib.CiReleaseCompaction(pq)
            self._lib.CiReleaseCompaction(pt)
            self._refresh_dims()

    def _map_ids(self, fn, ids):
        arr = _i64arr(ids).copy()
        if not fn(self.c_engine, arr.size, _p(arr, _pi64)):
            raise RuntimeError("id mapping failed")
        return arr

    def question_perm_from_comp(self, ids):
        return self._map_ids(self._lib.PqaEngine_QuestionPermFromComp, ids)

    def question_comp_from_perm(self, ids):
        return self._map_ids(self._lib.PqaEngine_QuestionCompFromPerm, ids)

    def target_perm_from_comp(self, ids):
        return self._map_ids(self._lib.PqaEngine_TargetPermFromComp, ids)

    def target_comp_from_perm(self, ids):
        return self._map_ids(self._lib.PqaEngine_TargetCompFromPerm, ids)

    def quiz_perm_from_comp(self, ids):
        return self._map_ids(self._lib.PqaEngine_QuizPermFromComp, ids)

    def quiz_comp_from_perm(self, ids):
        return self._map_ids(self._lib.PqaEngine_QuizCompFromPerm, ids)

    def save_kb_shard(self, file_path: str, write_frame: bool, throw: bool = True):
        """Sharded engines: this shard's cells into a file shared by all shards (the frame writer goes first)."""
        return _raise_or_return(self._lib.PqaB200_SaveKBShard(self.c_engine, file_path.encode(), 1 if write_frame else 0), throw)

    def shutdown(self, save_file_path: str = None, throw: bool = True):
        p = save_file_path.encode() if save_file_path else None
        return _raise_or_return(self._lib.PqaEngine_Shutdown(self.c_engine, p), throw)

    # ---------------------------------------------------------------- B200 extensions (numpy in / numpy out)
    @property
    def emulated_workers(self) -> int:
        return self._lib.PqaB200_GetEmulatedWorkers(self.c_engine)

    def upload_kb(self, sA, mD, vB):
        sA = np.ascontiguousarray(sA, dtype=np.float64)
        mD = np.ascontiguousarray(mD, dtype=np.float64)
        vB = np.ascontiguousarray(vB, dtype=np.float64)
        assert sA.shape == (self.n_questions, self.n_answers, self.n_targets), sA.shape
        assert mD.shape == (self.n_questions, self.n_targets) and vB.shape == (self.n_targets,)
        _raise_or_return(self._lib.PqaB200_UploadKB(self.c_engine, _p(sA, _pd), _p(mD, _pd), _p(vB, _pd)))

    def download_kb(self):
        sA = np.empty((self.n_questions, self.n_answers, self.n_targets))
        mD = np.empty((self.n_questions, self.n_targets))
        vB = np.empty(self.n_targets)
        _raise_or_return(self._lib.PqaB200_DownloadKB(self.c_engine, _p(sA, _pd), _p(mD, _pd), _p(vB, _pd)))
        return sA, mD, vB

    def copy_a_targets(self, i_question, i_answer):
        out = np.empty(self.n_targets)
        _raise_or_return(self._lib.PqaEngine_CopyATargets(self.c_engine, i_question, i_answer, self.n_targets, _p(out, _pd)))
        return out

    def copy_d_targets(self, i_question):
        out = np.empty(self.n_targets)
        _raise_or_return(self._lib.PqaEngine_CopyDTargets(self.c_engine, i_question, self.n_targets, _p(out, _pd)))
        return out

    def copy_b_targets(self):
        out = np.empty(self.n_targets)
        _raise_or_return(self._lib.PqaEngine_CopyBTargets(self.c_engine, self.n_targets, _p(out, _pd)))
        return out

    def start_quiz_batch(self, n: int) -> np.ndarray:
        ids = np.empty(n, dtype=np.int64)
        _raise_or_return(self._lib.PqaEngine_StartQuizBatch(self.c_engine, n, _p(ids, _pi64)))
        return ids

    def resume_quiz_batch(self, answered_lists) -> np.ndarray:
        """answered_lists: one list of (question, answer) pairs (or AnsweredQuestion) per quiz."""
        counts = np.array([len(l) for l in answered_lists], dtype=np.int64)
        flat = [aq for l in answered_lists for aq in l]
        arr, _ = self.to_c_answered_questions(flat)
        ids = np.empty(counts.size, dtype=np.int64)
        _raise_or_return(self._lib.PqaEngine_ResumeQuizBatch(self.c_engine, counts.size, _p(counts, _pi64), arr, _p(ids, _pi64)))
        return ids

    def next_question_batch(self, quiz_ids, randoms=None) -> np.ndarray:
        ids = _i64arr(quiz_ids)
        rnd = None if randoms is None else np.ascontiguousarray(randoms, dtype=np.uint64)
        out = np.empty(ids.size, dtype=np.int64)
        _raise_or_return(self._lib.PqaEngine_NextQuestionBatch(self.c_engine, ids.size, _p(ids, _pi64), _p(rnd, _pu64),
                                                               _p(out, _pi64), None))
        return out

    def record_answer_batch(self, quiz_ids, answers):
        ids, ans = _i64arr(quiz_ids), _i64arr(answers)
        assert ids.size == ans.size
        _raise_or_return(self._lib.PqaEngine_RecordAnswerBatch(self.c_engine, ids.size, _p(ids, _pi64), _p(ans, _pi64)))

    def set_active_question_batch(self, quiz_ids, questions):
        ids, qs = _i64arr(quiz_ids), _i64arr(questions)
        assert ids.size == qs.size
        _raise_or_return(self._lib.PqaEngine_SetActiveQuestionBatch(self.c_engine, ids.size, _p(ids, _pi64), _p(qs, _pi64)))

    def list_top_targets_batch(self, quiz_ids, max_count: int):
        """Returns (items[n, max_count] structured array {iTarget, prob}, counts[n])."""
        ids = _i64arr(quiz_ids)
        dest = np.zeros((ids.size, max_count), dtype=RATED_DTYPE)
        counts = np.zeros(ids.size, dtype=np.int64)
        _raise_or_return(self._lib.PqaEngine_ListTopTargetsBatch(self.c_engine, ids.size, _p(ids, _pi64), max_count,
                                                                 dest.ctypes.data_as(C.c_void_p), _p(counts, _pi64)))
        return dest, counts

    def record_quiz_target_batch(self, quiz_ids, targets, amounts=None):
        ids, tg = _i64arr(quiz_ids), _i64arr(targets)
        am = None if amounts is None else np.ascontiguousarray(amounts, dtype=np.float64)
        _raise_or_return(self._lib.PqaEngine_RecordQuizTargetBatch(self.c_engine, ids.size, _p(ids, _pi64), _p(tg, _pi64),
                                                                   _p(am, _pd)))

    def release_quiz_batch(self, quiz_ids):
        ids = _i64arr(quiz_ids)
        _raise_or_return(self._lib.PqaEngine_ReleaseQuizBatch(self.c_engine, ids.size, _p(ids, _pi64)))

    def copy_quiz_priors(self, i_quiz: int) -> np.ndarray:
        out = np.empty(self.n_targets)
        _raise_or_return(self._lib.PqaB200_CopyQuizPriors(self.c_engine, i_quiz, _p(out, _pd)))
        return out

    def set_quiz_priors(self, i_quiz: int, priors):
        pr = np.ascontiguousarray(priors, dtype=np.float64)
        assert pr.shape == (self.n_targets,)
        _raise_or_return(self._lib.PqaB200_SetQuizPriors(self.c_engine, i_quiz, _p(pr, _pd)))

    def eval_questions(self, quiz_ids):
        """dict(priority[n,Q] (NaN where asked), runLength[n,Q], grand[n,nChunks])."""
        ids = _i64arr(quiz_ids)
        n, Q = ids.size, self.n_questions
        pri, run = np.empty((n, Q)), np.empty((n, Q))
        maxChunks = min(Q, 8 * self.emulated_workers)
        grand = np.empty((n, maxChunks))
        nch = C.c_int64()
        _raise_or_return(self._lib.PqaB200_EvalQuestions(self.c_engine, n, _p(ids, _pi64), _p(pri, _pd), _p(run, _pd),
                                                         _p(grand, _pd), C.byref(nch)))
        assert nch.value == maxChunks
        return dict(priority=pri, runLength=run, grand=grand)

    def eval_questions_detailed(self, i_quiz: int):
        Q, K = self.n_questions, self.n_answers
        W, H, V = np.empty((Q, K)), np.empty((Q, K)), np.empty((Q, K))
        lack, pri = np.empty(Q), np.empty(Q)
        _raise_or_return(self._lib.PqaB200_EvalQuestionsDetailed(self.c_engine, i_quiz, _p(W, _pd), _p(H, _pd), _p(V, _pd),
                                                                 _p(lack, _pd), _p(pri, _pd)))
        return dict(W=W, H=H, V=V, lack=lack, priority=pri)

    def eval_questions_detailed_batch(self, quiz_ids):
        """dict(W, H, V [n,Q,K], lack, priority [n,Q]) from the kernel a NextQuestion batch of this size runs on."""
        ids = _i64arr(quiz_ids)
        n, Q, K = ids.size, self.n_questions, self.n_answers
        W, H, V = np.empty((n, Q, K)), np.empty((n, Q, K)), np.empty((n, Q, K))
        lack, pri = np.empty((n, Q)), np.empty((n, Q))
        _raise_or_return(self._lib.PqaB200_EvalQuestionsDetailedBatch(self.c_engine, n, _p(ids, _pi64), _p(W, _pd), _p(H, _pd),
                                                                      _p(V, _pd), _p(lack, _pd), _p(pri, _pd)))
        return dict(W=W, H=H, V=V, lack=lack, priority=pri)

    def set_eval_kernel(self, which: int, chunk_targets: int = 0, quizzes_per_cta: int = 0, kahan_lanes_per_thread: int = 0):
        """0 auto, 1 exact (CpuEngine rounding order), 2 staged (throughput kernel)."""
        _raise_or_return(self._lib.PqaB200_SetEvalTuning(self.c_engine, which, chunk_targets, quizzes_per_cta,
                                                         kahan_lanes_per_thread))

    # ---------------------------------------------------------------- question-sharded protocol (PqaB200Ext.h)
    def question_shard(self) -> Tuple[int, int]:
        first, count = C.c_int64(), C.c_int64()
        _raise_or_return(self._lib.PqaB200_GetQuestionShard(self.c_engine, C.byref(first), C.byref(count)))
        return first.value, count.value

    def shard_buffer(self, which: int) -> Tuple[int, int]:
        """(device pointer, number of doubles) of buffer 0 (priorities [n][Q]) or 1 (priors [n][Tp])."""
        ptr, cnt = C.c_void_p(), C.c_int64()
        _raise_or_return(self._lib.PqaB200_ShardBuffer(self.c_engine, which, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def shard_eval(self, quiz_ids):
        ids = _i64arr(quiz_ids)
        _raise_or_return(self._lib.PqaB200_ShardEval(self.c_engine, ids.size, _p(ids, _pi64)))

    def shard_select(self, quiz_ids, randoms) -> np.ndarray:
        ids = _i64arr(quiz_ids)
        rnd = np.ascontiguousarray(randoms, dtype=np.uint64)
        out = np.empty(ids.size, dtype=np.int64)
        _raise_or_return(self._lib.PqaB200_ShardSelect(self.c_engine, ids.size, _p(ids, _pi64), _p(rnd, _pu64),
                                                       _p(out, _pi64), None))
        return out

    def shard_record_answer_begin(self, quiz_ids, answers):
        ids, ans = _i64arr(quiz_ids), _i64arr(answers)
        _raise_or_return(self._lib.PqaB200_ShardRecordAnswerBegin(self.c_engine, ids.size, _p(ids, _pi64), _p(ans, _pi64)))

    def shard_record_answer_end(self, quiz_ids):
        ids = _i64arr(quiz_ids)
        _raise_or_return(self._lib.PqaB200_ShardRecordAnswerEnd(self.c_engine, ids.size, _p(ids, _pi64)))

    # ---------------------------------------------------------------- target-sharded protocol (PqaB200Ext.h)
    def target_shard(self) -> Tuple[int, int]:
        first, count = C.c_int64(), C.c_int64()
        _raise_or_return(self._lib.PqaB200_GetTargetShard(self.c_engine, C.byref(first), C.byref(count)))
        return first.value, count.value

    def tshard_eval_w(self, quiz_ids):
        ids = _i64arr(quiz_ids)
        _raise_or_return(self._lib.PqaB200_TShardEvalW(self.c_engine, ids.size, _p(ids, _pi64)))

    def tshard_eval_hvl(self, quiz_ids):
        ids = _i64arr(quiz_ids)
        _raise_or_return(self._lib.PqaB200_TShardEvalHVL(self.c_engine, ids.size, _p(ids, _pi64)))

    def tshard_priority(self, quiz_ids):
        ids = _i64arr(quiz_ids)
        _raise_or_return(self._lib.PqaB200_TShardPriority(self.c_engine, ids.size, _p(ids, _pi64)))

    def fill_binary_search_kb(self, rounds: float = 3.0):
        """Device-side fill of this engine's shard with synth.binary_search_kb(Q, K, T, init_amount, rounds)."""
        _raise_or_return(self._lib.PqaB200_FillBinarySearchKB(self.c_engine, float(rounds)))

    # ---------------------------------------------------------------- shard exchange over peer memory (PqaB200Ext.h)
    def p2p_init(self, rank: int, n_ranks: int, max_quizzes: int) -> Tuple[int, int]:
        """Allocates this shard's inbox; returns (device pointer, bytes)."""
        base, nbytes = C.c_void_p(), C.c_int64()
        _raise_or_return(self._lib.PqaB200_P2PInit(self.c_engine, rank, n_ranks, max_quizzes, C.byref(base), C.byref(nbytes)))
        return base.value, nbytes.value

    def p2p_export_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _raise_or_return(self._lib.PqaB200_P2PExportHandle(self.c_engine, buf))
        return buf.raw

    def p2p_open_handle(self, handle: bytes) -> int:
        base = C.c_void_p()
        _raise_or_return(self._lib.PqaB200_P2POpenHandle(self.c_engine, C.create_string_buffer(handle, 64), C.byref(base)))
        return base.value

    def p2p_connect(self, bases):
        arr = (C.c_void_p * len(bases))(*[C.c_void_p(b) for b in bases])
        _raise_or_return(self._lib.PqaB200_P2PConnect(self.c_engine, arr))

    def p2p_next_question_begin(self, quiz_ids, randoms):
        ids = _i64arr(quiz_ids)
        rnd = np.ascontiguousarray(randoms, dtype=np.uint64)
        _raise_or_return(self._lib.PqaB200_P2PNextQuestionBegin(self.c_engine, ids.size, _p(ids, _pi64), _p(rnd, _pu64)))

    def p2p_next_question_end(self, quiz_ids) -> np.ndarray:
        ids = _i64arr(quiz_ids)
        out = np.empty(ids.size, dtype=np.int64)
        _raise_or_return(self._lib.PqaB200_P2PNextQuestionEnd(self.c_engine, ids.size, _p(ids, _pi64), _p(out, _pi64), None))
        return out

    def p2p_record_answer_begin(self, quiz_ids, answers):
        ids, ans = _i64arr(quiz_ids), _i64arr(answers)
        _raise_or_return(self._lib.PqaB200_P2PRecordAnswerBegin(self.c_engine, ids.size, _p(ids, _pi64), _p(ans, _pi64)))

    def anomaly_counts(self) -> np.ndarray:
        """[priorities <= 0 or non-finite, non-finite running totals, grand totals <= 0] seen by NextQuestion so far (the
        reference logs warnings for these and carries on: CEEvalQsSubtaskConsider.cpp:209-211, CpuEngine.cpp:368-377)."""
        out = np.zeros(3, dtype=np.uint64)
        _raise_or_return(self._lib.PqaB200_AnomalyCounts(self.c_engine, _p(out, _pu64)))
        return out

    def p2p_last_phase_ms(self) -> np.ndarray:
        """[phase 1, barrier, phase 2, barrier, epilogue + selection] device ms of the last target-sharded P2PNextQuestion."""
        out = np.zeros(5)
        _raise_or_return(self._lib.PqaB200_P2PLastPhaseMs(self.c_engine, _p(out, _pd)))
        return out

    def p2p_set_exact_order(self, on: bool):
        """Target shards: hand the Kahan lanes from shard to shard so that W_k is bit-identical to a single engine's."""
        _raise_or_return(self._lib.PqaB200_P2PSetExactOrder(self.c_engine, 1 if on else 0))

    def p2p_record_answer_end(self):
        _raise_or_return(self._lib.PqaB200_P2PRecordAnswerEnd(self.c_engine))

    def resident_bind(self, quiz_ids, randoms=None):
        ids = _i64arr(quiz_ids)
        rnd = None if randoms is None else np.ascontiguousarray(randoms, dtype=np.uint64)
        _raise_or_return(self._lib.PqaB200_ResidentBind(self.c_engine, ids.size, _p(ids, _pi64), _p(rnd, _pu64)))
        self._resident_n = ids.size

    def resident_step(self):
        _raise_or_return(self._lib.PqaB200_ResidentStep(self.c_engine))

    def resident_fetch(self) -> np.ndarray:
        out = np.empty(self._resident_n, dtype=np.int64)
        _raise_or_return(self._lib.PqaB200_ResidentFetch(self.c_engine, _p(out, _pi64)))
        return out

    def resident_last_eval_ms(self) -> float:
        return self._lib.PqaB200_ResidentLastEvalMs(self.c_engine)

    def synchronize(self):
        _raise_or_return(self._lib.PqaB200_Synchronize(self.c_engine))

    def flush_l2(self):
        _raise_or_return(self._lib.PqaB200_FlushL2(self.c_engine))

    def shard_count(self) -> int:
        return int(self._lib.PqaB200_GetShardCount(self.c_engine))

    def kernel_launch_count(self) -> int:
        return self._lib.PqaB200_KernelLaunchCount(self.c_engine)


class DeviceEvent:
    """CUDA event recorded on the engine's own stream (torch.cuda.Event would only see torch's current stream)."""

    def __init__(self):
        self._lib = load_library()
        self.h = C.c_void_p(self._lib.PqaB200_EventCreate())
        if not self.h.value:
            raise RuntimeError("cudaEventCreate failed")

    def record(self, engine: PqaEngine):
        _raise_or_return(self._lib.PqaB200_EventRecord(engine.c_engine, self.h))

    def resident_last_eval_ms(self) -> float:
        return self._lib.PqaB200_ResidentLastEvalMs(self.c_engine)

    def synchronize(self):
        _raise_or_return(self._lib.PqaB200_EventSynchronize(self.h))

    def elapsed_ms(self, stop: "DeviceEvent") -> float:
        return self._lib.PqaB200_EventElapsedMs(self.h, stop.h)

    def __del__(self):
        if getattr(self, "h", None) and self.h.value:
            self._lib.PqaB200_EventDestroy(self.h)
            self.h = C.c_void_p(None)


class PqaEngineFactory:
    def __init__(self):
        self._lib = load_library()
        self.c_factory = C.c_void_p(self._lib.CiGetPqaEngineFactory())

    def create_cpu_engine(self, eng_def: EngineDefinition) -> Tuple[PqaEngine, PqaError]:
        """Same name and shape as ProbQA.py:747; returns the B200 engine (there is no CPU engine in this library)."""
        c_def = eng_def.to_c()
        e = C.c_void_p()
        c_engine = self._lib.PqaEngineFactory_CreateCpuEngine(self.c_factory, C.byref(e), C.byref(c_def))
        err = PqaError.factor(e.value)
        if not c_engine:
            raise PqaException(err)
        return PqaEngine(c_engine), err

    def create_b200_engine(self, eng_def: EngineDefinition, device: int = -1, emulated_workers: int = 0,
                           rng_seed: int = 0, initial_quiz_capacity: int = 0, question_shard_first: int = 0,
                           question_shard_count: int = 0, target_shard_first: int = 0,
                           target_shard_count: int = 0) -> PqaEngine:
        c_def = eng_def.to_c()
        opts = CiB200Options(device, emulated_workers, rng_seed, initial_quiz_capacity, question_shard_first,
                             question_shard_count, target_shard_first, target_shard_count)
        e = C.c_void_p()
        c_engine = self._lib.PqaB200_CreateEngine(C.byref(e), C.byref(c_def), C.byref(opts))
        if not c_engine:
            raise PqaException(PqaError.factor(e.value))
        return PqaEngine(c_engine)

    @staticmethod
    def _group_options(axis, n_shards, devices, exact_order, max_batch):
        g = CiB200GroupOptions()
        g.axis = {"questions": 0, "targets": 1}[axis]
        g.nShards = n_shards
        for r in range(8):
            g.devices[r] = devices[r] if devices is not None and r < len(devices) else -1
        g.exactOrder = 1 if exact_order else 0
        g.maxBatch = max_batch
        return g

    def create_sharded_engine(self, eng_def: EngineDefinition, axis: str, n_shards: int, devices=None, exact_order: bool = False,
                              max_batch: int = 0, emulated_workers: int = 0, rng_seed: int = 0, initial_quiz_capacity: int = 0) -> PqaEngine:
        """One engine handle over n_shards shard engines of THIS process (one per listed device; several shards may share
        a device), exchanging over peer memory. The handle serves the reference ABI and the batch calls like one engine."""
        c_def = eng_def.to_c()
        opts = CiB200Options(-1, emulated_workers, rng_seed, initial_quiz_capacity, 0, 0, 0, 0)
        g = self._group_options(axis, n_shards, devices, exact_order, max_batch)
        e = C.c_void_p()
        c_engine = self._lib.PqaB200_CreateShardedEngine(C.byref(e), C.byref(c_def), C.byref(opts), C.byref(g))
        if not c_engine:
            raise PqaException(PqaError.factor(e.value))
        return PqaEngine(c_engine)

    def load_sharded_engine(self, file_path: str, axis: str, n_shards: int, devices=None, exact_order: bool = False,
                            max_batch: int = 0, emulated_workers: int = 0, rng_seed: int = 0) -> PqaEngine:
        opts = CiB200Options(-1, emulated_workers, rng_seed, 0, 0, 0, 0, 0)
        g = self._group_options(axis, n_shards, devices, exact_order, max_batch)
        e = C.c_void_p()
        c_engine = self._lib.PqaB200_LoadShardedEngine(C.byref(e), file_path.encode(), C.byref(opts), C.byref(g))
        if not c_engine:
            raise PqaException(PqaError.factor(e.value))
        return PqaEngine(c_engine)

    def load_b200_engine(self, file_path: str, device: int = -1, emulated_workers: int = 0, rng_seed: int = 0,
                         initial_quiz_capacity: int = 0, question_shard_first: int = 0, question_shard_count: int = 0,
                         target_shard_first: int = 0, target_shard_count: int = 0) -> PqaEngine:
        """LoadCpuEngine with explicit options; a sharded engine reads only its rows / columns of the file."""
        opts = CiB200Options(device, emulated_workers, rng_seed, initial_quiz_capacity, question_shard_first,
                             question_shard_count, target_shard_first, target_shard_count)
        e = C.c_void_p()
        c_engine = self._lib.PqaB200_LoadEngine(C.byref(e), file_path.encode(), C.byref(opts))
        if not c_engine:
            raise PqaException(PqaError.factor(e.value))
        return PqaEngine(c_engine)

    def load_cpu_engine(self, file_path: str, mem_pool_max_bytes: int = 512 << 20) -> Tuple[PqaEngine, PqaError]:
        e = C.c_void_p()
        c_engine = self._lib.PqaEngineFactory_LoadCpuEngine(self.c_factory, C.byref(e), file_path.encode(), mem_pool_max_bytes)
        err = PqaError.factor(e.value)
        if not c_engine:
            raise PqaException(err)
        return PqaEngine(c_engine), err
