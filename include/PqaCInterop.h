/* probqa_b200 -- C ABI of libPqaCore.so (B200 / sm_100a engine).
 *
 * Drop-in boundary: every symbol, signature and struct layout below mirrors the reference's
 * ProbQA/PqaCore/Interface/PqaCInterop.h:9-108 (structs :9-42 under #pragma pack(push, 8), functions :48-108).
 * A client of the reference's PqaCore.dll (the ctypes binding Interop/Python/ProbQAInterop/ProbQA.py:72-294, the
 * .NET P/Invoke wrapper ProbQANetCore/PqaEngine.cs, a C++ caller of the extern "C" layer) binds to this library
 * unchanged. Reference-side definitions of each entry point: ProbQA/PqaCore/PqaCInterop.cpp (line cited per symbol).
 *
 * Error convention (PqaCInterop.cpp:45-61): functions returning void* return NULL on success or an owned error
 * object; functions returning a value take void **ppError (set to NULL or an owned error) and return -1 / 0 on
 * failure. Release errors with CiReleasePqaError, stringify with PqaError_ToString (string released by
 * CiReleaseString).
 *
 * Additive B200 entry points (batches of concurrent quizzes, KB bulk transfer, device timing) are declared in
 * PqaB200Ext.h; nothing here depends on them. */
#ifndef PQA_C_INTEROP_H
#define PQA_C_INTEROP_H

#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#if defined(_WIN32)
#define PQACORE_API __declspec(dllimport)
#else
#define PQACORE_API __attribute__((visibility("default")))
#endif

#pragma pack(push, 8)
typedef struct {           /* PqaCInterop.h:10-19 */
  int64_t _nAnswers;
  int64_t _nQuestions;
  int64_t _nTargets;
  uint8_t _precType;       /* TPqaPrecisionType, Interface/PqaCommon.h:17-24: 1 Float, 3 Double (only Double accepted) */
  uint16_t _precExponent;
  uint32_t _precMantissa;
  double _initAmount;
  uint64_t _memPoolMaxBytes;
} CiEngineDefinition;

typedef struct {           /* PqaCInterop.h:21-24 */
  int64_t _iQuestion;
  int64_t _iAnswer;
} CiAnsweredQuestion;

typedef struct {           /* PqaCInterop.h:26-30 */
  int64_t _nAnswers;
  int64_t _nQuestions;
  int64_t _nTargets;
} CiEngineDimensions;

typedef struct {           /* PqaCInterop.h:32-35 */
  int64_t _iTarget;
  double _prob;
} CiRatedTarget;

typedef struct {           /* PqaCInterop.h:37-40 */
  int64_t _index;
  double _initAmount;
} CiAddQorTParam;
#pragma pack(pop)

#ifdef __cplusplus
extern "C" {
#endif

PQACORE_API void CiDebugBreak(void);                                                          /* PqaCInterop.cpp:314-316 */

PQACORE_API uint8_t Logger_Init(void **ppStrErr, const char *baseName);                       /* :145-172 */
PQACORE_API void CiReleaseString(void *pvString);                                             /* :134-137 */

PQACORE_API void *CiGetPqaEngineFactory(void);                                                /* :88-90 */
/* Creates the B200 engine (there is no CPU fallback); precType must be Double like the reference's CPU factory
 * (PqaEngineBaseFactory.cpp:16-27); dims >= 2 answers, 1 question, 2 targets (PqaEngineBaseFactory.h:15-17). */
PQACORE_API void *PqaEngineFactory_CreateCpuEngine(void *pvFactory, void **ppError,
                                                   const CiEngineDefinition *pEngDef);       /* :92-112 */
PQACORE_API void *PqaEngineFactory_LoadCpuEngine(void *pvFactory, void **ppError, const char *filePath,
                                                 uint64_t memPoolMaxBytes);                   /* :114-127 */

PQACORE_API void CiReleasePqaError(void *pvErr);                                              /* :129-132 */
PQACORE_API void *PqaError_ToString(void *pvError, const uint8_t withParams);                 /* :139-143 */

PQACORE_API void CiReleasePqaEngine(void *pvEngine);                                          /* :174-177 */
PQACORE_API void *PqaEngine_Train(void *pvEngine, int64_t nQuestions, const CiAnsweredQuestion *const pAQs,
                                  const int64_t iTarget, const double amount);                /* :179-184 */

PQACORE_API uint8_t PqaEngine_QuestionPermFromComp(void *pvEngine, const int64_t count, int64_t *pIds); /* :186-189 */
PQACORE_API uint8_t PqaEngine_QuestionCompFromPerm(void *pvEngine, const int64_t count, int64_t *pIds); /* :191-194 */
PQACORE_API uint8_t PqaEngine_TargetPermFromComp(void *pvEngine, const int64_t count, int64_t *pIds);   /* :196-199 */
PQACORE_API uint8_t PqaEngine_TargetCompFromPerm(void *pvEngine, const int64_t count, int64_t *pIds);   /* :201-204 */
PQACORE_API uint8_t PqaEngine_QuizPermFromComp(void *pvEngine, const int64_t count, int64_t *pIds);     /* :206-209 */
PQACORE_API uint8_t PqaEngine_QuizCompFromPerm(void *pvEngine, const int64_t count, int64_t *pIds);     /* :211-214 */
PQACORE_API uint8_t PqaEngine_EnsurePermQuizGreater(void *pvEngine, const int64_t bound);               /* :216-219 */
PQACORE_API uint8_t PqaEngine_RemapQuizPermId(void *pvEngine, const int64_t srcPermId, const int64_t destPermId); /* :221-224 */

PQACORE_API uint64_t PqaEngine_GetTotalQuestionsAsked(void *pvEngine, void **ppError);        /* :226-232 */
PQACORE_API uint8_t PqaEngine_CopyDims(void *pvEngine, CiEngineDimensions *pDims);            /* :234-241 */
PQACORE_API int64_t PqaEngine_StartQuiz(void *pvEngine, void **ppError);                      /* :243-249 */
PQACORE_API int64_t PqaEngine_ResumeQuiz(void *pvEngine, void **ppError, const int64_t nAnswered,
                                         const CiAnsweredQuestion *const pAQs);               /* :251-259 */
PQACORE_API int64_t PqaEngine_NextQuestion(void *pvEngine, void **ppError, const int64_t iQuiz); /* :261-267 */
PQACORE_API void *PqaEngine_RecordAnswer(void *pvEngine, const int64_t iQuiz, const int64_t iAnswer); /* :269-272 */

PQACORE_API void *PqaEngine_ClearOldQuizzes(void *pvEngine, const int64_t maxCount, const double maxAgeSec); /* :379-382 */

PQACORE_API int64_t PqaEngine_GetActiveQuestionId(void *pvEngine, void **ppError, const int64_t iQuiz); /* :301-307 */
PQACORE_API void *PqaEngine_SetActiveQuestion(void *pvEngine, const int64_t iQuiz, const int64_t iQuestion); /* :309-312 */

PQACORE_API int64_t PqaEngine_ListTopTargets(void *pvEngine, void **ppError, const int64_t iQuiz,
                                             const int64_t maxCount, CiRatedTarget *pDest);   /* :274-282 */
PQACORE_API void *PqaEngine_RecordQuizTarget(void *pvEngine, const int64_t iQuiz, const int64_t iTarget,
                                             const double amount);                            /* :284-289 */
PQACORE_API void *PqaEngine_ReleaseQuiz(void *pvEngine, const int64_t iQuiz);                 /* :291-294 */
PQACORE_API void *PqaEngine_SaveKB(void *pvEngine, const char *const filePath, const uint8_t bDoubleBuffer); /* :296-299 */

PQACORE_API void *PqaEngine_StartMaintenance(void *pvEngine, const bool forceQuizzes);        /* :318-321 */
PQACORE_API void *PqaEngine_FinishMaintenance(void *pvEngine);                                /* :323-326 */
PQACORE_API void *PqaEngine_AddQsTs(void *pvEngine, const int64_t nQuestions, CiAddQorTParam *pAddQuestionParams,
                                    const int64_t nTargets, CiAddQorTParam *pAddTargetParams); /* :328-334 */
PQACORE_API void *PqaEngine_RemoveQuestions(void *pvEngine, const int64_t nQuestions, const int64_t *pQIds); /* :336-339 */
PQACORE_API void *PqaEngine_RemoveTargets(void *pvEngine, const int64_t nTargets, const int64_t *pTIds);     /* :341-344 */
PQACORE_API void *PqaEngine_Compact(void *pvEngine, int64_t *pnQuestions, int64_t const **const ppOldQuestions,
                                    int64_t *pnTargets, int64_t const **const ppOldTargets);  /* :346-362 */
PQACORE_API void CiReleaseCompaction(const int64_t *p);                                       /* :364-366 */
PQACORE_API void *PqaEngine_Shutdown(void *pvEngine, const char *const saveFilePath);         /* :368-371 */
PQACORE_API void *PqaEngine_SetLogger(void *pvEngine, void *pSRLogger);                       /* :373-377 */

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* PQA_C_INTEROP_H */
