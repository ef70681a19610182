/* probqa_b200 -- additive C entry points of libPqaCore.so that the reference's one-quiz-per-call ABI cannot express.
 *
 * None of these exists in the reference (ProbQA/PqaCore/Interface/PqaCInterop.h); they follow its conventions
 * (void* = NULL or owned PqaError; plain pointers and sizes; caller owns every buffer). They exist because
 * BASELINE configs 2-4 are batches of concurrent quizzes sharing one stream over sA/mD (SURVEY.md 8b last row).
 * All host pointers are ordinary host memory unless the parameter name starts with "d" (device pointer). */
#ifndef PQA_B200_EXT_H
#define PQA_B200_EXT_H
#include "PqaCInterop.h"

#pragma pack(push, 8)
typedef struct {
  int32_t _device;           /* CUDA device ordinal; -1 = current device / env PQA_B200_DEVICE */
  /* Worker count W of the CpuEngine being reproduced (reference: std::thread::hardware_concurrency(),
   * BaseCpuEngine.cpp:21-22). It fixes the order of the Kahan sums in StartQuiz/RecordAnswer, the run-length
   * chunking of NextQuestion and the piece split of ListTopTargets, i.e. the bits of the results.
   * 0 = env PQA_B200_EMULATED_WORKERS, else this host's hardware_concurrency(). */
  int32_t _emulatedWorkers;
  uint64_t _rngSeed;         /* seed of the host xorshift128+ used by NextQuestion; 0 = std::random_device */
  int64_t _initialQuizCapacity; /* quiz slots pre-allocated on the device; 0 = 256 */
  /* Question shard of a multi-GPU engine: this device holds the sA/mD rows of questions
   * [_questionShardFirst, _questionShardFirst + _questionShardCount) only; _questionShardCount = 0 means all questions.
   * Quiz state (priors, asked bits) is replicated on every shard; see the "question-sharded" entry points below. */
  int64_t _questionShardFirst;
  int64_t _questionShardCount;
  /* Target shard of a multi-GPU engine (BASELINE config 4): this device holds columns
   * [_targetShardFirst, _targetShardFirst + _targetShardCount) of every sA/mD row; _targetShardCount = 0 means all
   * targets. First and count must be multiples of 4 targets (the last shard ends at nTargets). Quiz state is replicated
   * with full-length priors; see the "target-sharded" entry points below. Not combinable with a question shard. */
  int64_t _targetShardFirst;
  int64_t _targetShardCount;
} CiB200Options;
/* One engine over several GPUs of a box in ONE process (PqaB200_CreateShardedEngine): the KB is split over _nShards shard
 * engines, their kernels exchange over NVLink peer memory, and the returned handle answers every entry point of
 * PqaCInterop.h and the batch entry points below like a single engine does. */
typedef struct {
  int32_t _axis;             /* 0 = questions (rows of sA/mD; results bit-identical to one engine), 1 = targets (columns;
                              * BASELINE config 4) */
  int32_t _nShards;          /* 1..8 */
  int32_t _devices[8];       /* CUDA device of each shard; all -1 = devices 0.._nShards-1 modulo the visible device count */
  int32_t _exactOrder;       /* targets only: hand the Kahan lanes from shard to shard (PqaB200_P2PSetExactOrder) */
  int64_t _maxBatch;         /* quizzes per exchanged call (inbox size); 0 = 256; longer batches are cut into slices */
} CiB200GroupOptions;
#pragma pack(pop)

#ifdef __cplusplus
extern "C" {
#endif

/* A sharded engine group behind the ordinary engine handle. pOpts: as for PqaB200_CreateEngine (shard fields ignored).
 * Environment shortcut for unmodified clients of the reference ABI: with PQA_B200_SHARDS=N set,
 * PqaEngineFactory_CreateCpuEngine / LoadCpuEngine return such a group (PQA_B200_SHARD_AXIS=questions|targets, default
 * targets; PQA_B200_SHARD_DEVICES=0,1,..; PQA_B200_SHARD_EXACT=1). */
PQACORE_API void *PqaB200_CreateShardedEngine(void **ppError, const CiEngineDefinition *pEngDef, const CiB200Options *pOpts,
                                              const CiB200GroupOptions *pGroupOpts);
PQACORE_API void *PqaB200_LoadShardedEngine(void **ppError, const char *filePath, const CiB200Options *pOpts,
                                            const CiB200GroupOptions *pGroupOpts);
/* Number of shard engines behind the handle (1 for a plain engine). */
PQACORE_API int32_t PqaB200_GetShardCount(void *pvEngine);

/* Same contract as PqaEngineFactory_CreateCpuEngine, with explicit options. */
PQACORE_API void *PqaB200_CreateEngine(void **ppError, const CiEngineDefinition *pEngDef, const CiB200Options *pOpts);
/* Same contract as PqaEngineFactory_LoadCpuEngine, with explicit options: a sharded engine streams only its question rows /
 * target columns out of the file (the whole KB is never staged on the host). */
PQACORE_API void *PqaB200_LoadEngine(void **ppError, const char *filePath, const CiB200Options *pOpts);
/* Sharded engines save into ONE file in the reference's layout: the shard called with writeFrame != 0 goes first (creates
 * the file; header, vB, gap lists, id maps), then every other shard writes its cells in place (writeFrame = 0). */
PQACORE_API void *PqaB200_SaveKBShard(void *pvEngine, const char *filePath, int32_t writeFrame);
PQACORE_API int32_t PqaB200_GetEmulatedWorkers(void *pvEngine);
PQACORE_API int32_t PqaB200_GetDevice(void *pvEngine);
PQACORE_API const char *PqaB200_BuildInfo(void); /* static string: arch, build flags */
/* Self test of the host-side bookkeeping that needs no device (gap sets, permanent ids, compaction plans): returns NULL
 * when all is well, else a message to release with CiReleaseString. */
PQACORE_API void *PqaB200_HostLogicSelfTest(void);

/* C exports of IPqaEngine::CopyATargets/CopyDTargets/CopyBTargets (Interface/IPqaEngine.h:36-39; C++-only in the
 * reference). */
PQACORE_API void *PqaEngine_CopyATargets(void *pvEngine, int64_t iQuestion, int64_t iAnswer, int64_t maxTargets, double *pFreqs);
PQACORE_API void *PqaEngine_CopyDTargets(void *pvEngine, int64_t iQuestion, int64_t maxTargets, double *pFreqs);
PQACORE_API void *PqaEngine_CopyBTargets(void *pvEngine, int64_t maxTargets, double *pFreqs);

/* Whole-KB transfer in the reference's file layout (CpuEngine.cpp:664-688): sA[(i*K+k)*T + j], mD[i*T + j], vB[j]. */
PQACORE_API void *PqaB200_UploadKB(void *pvEngine, const double *sA, const double *mD, const double *vB);
PQACORE_API void *PqaB200_DownloadKB(void *pvEngine, double *sA, double *mD, double *vB);

/* ---- batches of concurrent quizzes (each quiz id must appear at most once per call) ---- */
PQACORE_API void *PqaEngine_StartQuizBatch(void *pvEngine, int64_t n, int64_t *pQuizIds);
/* n quizzes resumed at once (PqaEngine_ResumeQuiz semantics each): quiz x has answered pCounts[x] questions, taken in
 * order from pAQs; pQuizIds[x] receives its id (-1 for a quiz that hit the reference's I64Underflow error). */
PQACORE_API void *PqaEngine_ResumeQuizBatch(void *pvEngine, int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs,
                                            int64_t *pQuizIds);
/* pRandoms: optional n 64-bit draws replacing the engine RNG (SRDoubleNumber.h:35-39 consumes one per call).
 * pQuestions[i] = chosen question or -1; ppErrors: optional n slots receiving NULL / owned per-quiz errors. */
PQACORE_API void *PqaEngine_NextQuestionBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds,
                                              const uint64_t *pRandoms, int64_t *pQuestions, void **ppErrors);
PQACORE_API void *PqaEngine_RecordAnswerBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers);
PQACORE_API void *PqaEngine_SetActiveQuestionBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pQuestions);
/* pDest: n*maxCount items, row i belongs to quiz i; pCounts[i] = number listed. */
PQACORE_API void *PqaEngine_ListTopTargetsBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, int64_t maxCount,
                                                CiRatedTarget *pDest, int64_t *pCounts);
/* Applied in array order, exactly as n successive PqaEngine_RecordQuizTarget calls would be. pAmounts may be NULL (=1). */
PQACORE_API void *PqaEngine_RecordQuizTargetBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds,
                                                  const int64_t *pTargets, const double *pAmounts);
PQACORE_API void *PqaEngine_ReleaseQuizBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds);

/* ---- inspection (parity tests, diagnostics) ---- */
PQACORE_API void *PqaB200_CopyQuizPriors(void *pvEngine, int64_t iQuiz, double *pPriors /* nTargets */);
PQACORE_API void *PqaB200_SetQuizPriors(void *pvEngine, int64_t iQuiz, const double *pPriors /* nTargets */);
/* Evaluates every question of each quiz without selecting one. Any output may be NULL.
 * pPriorities[n*Q] (NaN for asked questions), pRunLength[n*Q] (chunk-local Kahan prefixes, CpuEngine.cpp:337-360),
 * pGrandTotals[n*nChunks] with nChunks = min(Q, 8*W) returned in *pnChunks. */
PQACORE_API void *PqaB200_EvalQuestions(void *pvEngine, int64_t n, const int64_t *pQuizIds, double *pPriorities,
                                        double *pRunLength, double *pGrandTotals, int64_t *pnChunks);
/* Per-answer metrics of one quiz: pW/pH/pV [Q*K], pLack [Q] (CEEvalQsSubtaskConsider.cpp:88,129-132,201). */
PQACORE_API void *PqaB200_EvalQuestionsDetailed(void *pvEngine, int64_t iQuiz, double *pW, double *pH, double *pV,
                                                double *pLack, double *pPriorities);
/* The same for a batch of n quizzes, evaluated by the kernel a NextQuestion batch of that size runs on: pW/pH/pV [n*Q*K],
 * pLack / pPriorities [n*Q]. */
PQACORE_API void *PqaB200_EvalQuestionsDetailedBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, double *pW, double *pH,
                                                     double *pV, double *pLack, double *pPriorities);
/* Selects which evaluation kernel the engine uses: 0 auto (= 2), 1 exact (every rounding of CpuEngine reproduced;
 * bit-identical W/H/V/lack), 2 staged (TMA + shared memory throughput kernel, tolerance-level parity). */
PQACORE_API void *PqaB200_SetEvalKernel(void *pvEngine, int32_t which);
/* Same, plus tuning knobs of the staged kernel (tests force the chunked-targets path on small T with these):
 * chunkTargets = targets staged per shared-memory chunk (0 auto), quizzesPerCta = quiz tile of one CTA (0 auto),
 * kahanLanesPerThread = 4 (one thread per quiz), 2 (two threads per quiz), 1 (four threads per quiz) or 0 = auto:
 * batches > 64 two threads per quiz in 8-warp CTAs; 33..64 quizzes four threads per quiz (whole slab) or two threads per
 * quiz in 4-warp CTAs (chunked targets); fewer than 32 the small-batch kernel. */
PQACORE_API void *PqaB200_SetEvalTuning(void *pvEngine, int32_t which, int64_t chunkTargets, int64_t quizzesPerCta,
                                        int32_t kahanLanesPerThread);

/* ---- question-sharded engines (one engine per GPU, each with a CiB200Options question shard) ----
 * The exchange between shards is done by the caller (NCCL all-reduce through torch.distributed, or any other sum) on
 * the device buffers returned by PqaB200_ShardBuffer; adding the other shards' zeros is exact, so results are those of a
 * single engine. Protocol per NextQuestion:   ShardEval -> all-reduce(buffer 0, n*Q doubles) -> ShardSelect.
 * Protocol per RecordAnswer:                  ShardRecordAnswerBegin -> all-reduce(buffer 1, n*Tp doubles) -> ...End.
 * StartQuiz / ListTopTargets / SetActiveQuestion are local (replicated state). RecordQuizTarget / Train apply the
 * operations of owned questions only; vB is updated on every shard. */
PQACORE_API void *PqaB200_ShardEval(void *pvEngine, int64_t n, const int64_t *pQuizIds);
PQACORE_API void *PqaB200_ShardSelect(void *pvEngine, int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms,
                                      int64_t *pQuestions, void **ppErrors);
PQACORE_API void *PqaB200_ShardRecordAnswerBegin(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers);
PQACORE_API void *PqaB200_ShardRecordAnswerEnd(void *pvEngine, int64_t n, const int64_t *pQuizIds);
/* which: 0 = priorities [n][Q], 1 = priors [n][Tp], 2 = W_k partials [n][Q][K], 3 = H/V/lack partials [n][Q][2K+1]
 * (2 and 3: target-sharded engines). Returns the device pointer and the number of doubles of the last
 * Shard* call that filled it. */
PQACORE_API void *PqaB200_ShardBuffer(void *pvEngine, int32_t which, void **ppDevice, int64_t *pCount);
PQACORE_API void *PqaB200_GetQuestionShard(void *pvEngine, int64_t *pFirst, int64_t *pCount);

/* ---- target-sharded engines (one engine per GPU, each with a CiB200Options target shard) ----
 * The question evaluation needs the complete normaliser W_k before the entropy / lack / distance sums can be formed
 * (log2(lik/W_k) is in the lack term's denominator), so it runs in two phases with a sum over shards after each:
 *   TShardEvalW   -> sum(buffer 2: n*Q*K partial W_k)            -> TShardEvalHVL
 *                 -> sum(buffer 3: n*Q*(2K+1) partial H_k,V_k,L) -> TShardPriority (fills buffer 0) -> ShardSelect.
 * Sums over shards are in a different order than CpuEngine's single Kahan pass: priorities are tolerance-level
 * (DESIGN.md), while posteriors stay bit-exact: ShardRecordAnswerBegin fills this shard's columns of the
 * un-normalised row (zeros elsewhere) -> sum(buffer 1: n*Tp) -> ShardRecordAnswerEnd normalises the complete row in the
 * reference's order on every shard. StartQuiz / ListTopTargets / SetActiveQuestion are local. RecordQuizTarget / Train
 * update the cells of owned targets; vB is replicated and updated on every shard.
 * PqaB200_UploadKB / DownloadKB / CopyATargets / CopyDTargets take whole-KB host arrays and touch only this shard's
 * columns of them. */
PQACORE_API void *PqaB200_TShardEvalW(void *pvEngine, int64_t n, const int64_t *pQuizIds);
PQACORE_API void *PqaB200_TShardEvalHVL(void *pvEngine, int64_t n, const int64_t *pQuizIds);
PQACORE_API void *PqaB200_TShardPriority(void *pvEngine, int64_t n, const int64_t *pQuizIds);
PQACORE_API void *PqaB200_GetTargetShard(void *pvEngine, int64_t *pFirst, int64_t *pCount);
/* Fills this engine's shard of the KB with the closed-form "binary-search trained" KB of SURVEY.md 8d
 * (probqa_b200/synth.py binary_search_kb, bit-identical) on the device: KBs too large to stage through the host. */
PQACORE_API void *PqaB200_FillBinarySearchKB(void *pvEngine, double rounds);

/* ---- shard exchange over peer memory (NVLink P2P), for question-sharded and target-sharded engines ----
 * Instead of handing buffers to the caller for an all-reduce, the engines exchange directly: each engine owns an
 * "inbox" allocation; the evaluation / RecordAnswer kernels store what the other shards need straight into the other
 * shards' inboxes from their epilogues, and a one-CTA barrier kernel (release/acquire flags in the inboxes) orders the
 * phases on the device. No host round trip or separate collective launch sits between the phases of a call.
 *   setup : P2PInit on every engine (same nRanks / maxQuizzes) -> exchange the inbox bases: within one process pass
 *           *ppBase around; between processes P2PExportHandle (64-byte cudaIpcMemHandle_t) -> all-gather -> P2POpenHandle
 *           -> P2PConnect(bases of all ranks, own entry ignored).
 *   calls : every shard issues the same sequence of P2PNextQuestionBegin/End and P2PRecordAnswerBegin/End calls with
 *           the same quiz ids, draws and answers. Begin only enqueues (so one thread can drive several engines of one
 *           process: Begin on all, then End on all); End waits and returns the results. Results equal those of the
 *           caller-exchanged Shard* / TShard* protocol bit for bit (shard partials are summed in rank order).
 * A shard that does not show up within 20 s makes the barrier give up: End returns an Internal error. */
PQACORE_API void *PqaB200_P2PInit(void *pvEngine, int32_t rank, int32_t nRanks, int64_t maxQuizzes, void **ppBase, int64_t *pBytes);
PQACORE_API void *PqaB200_P2PExportHandle(void *pvEngine, uint8_t *pHandle64);
PQACORE_API void *PqaB200_P2POpenHandle(void *pvEngine, const uint8_t *pHandle64, void **ppPeerBase);
PQACORE_API void *PqaB200_P2PConnect(void *pvEngine, void *const *pBases);
PQACORE_API void *PqaB200_P2PNextQuestionBegin(void *pvEngine, int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms);
PQACORE_API void *PqaB200_P2PNextQuestionEnd(void *pvEngine, int64_t n, const int64_t *pQuizIds, int64_t *pQuestions, void **ppErrors);
PQACORE_API void *PqaB200_P2PRecordAnswerBegin(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers);
PQACORE_API void *PqaB200_P2PRecordAnswerEnd(void *pvEngine);
/* The reference's question evaluation only WARNS (through its logger) about a priority that is <= 0 or not finite
 * (CEEvalQsSubtaskConsider.cpp:209-211), about non-finite running totals (CpuEngine.cpp:368-371) and about a grand total
 * <= 0 (CpuEngine.cpp:375-377), and carries on; so does this engine. pCounts3 receives how often each of the three has
 * been seen by NextQuestion since the engine was created; a growing count is also reported on stderr. */
PQACORE_API void *PqaB200_AnomalyCounts(void *pvEngine, uint64_t *pCounts3);
/* Target shards: device time (ms, CUDA events on the engine's stream) of the five stages of the most recent
 * P2PNextQuestion -- phase 1, exchange barrier, phase 2, exchange barrier, epilogue + selection. Call after its End. */
PQACORE_API void *PqaB200_P2PLastPhaseMs(void *pvEngine, double *pMs5);
/* Target shards only. on != 0: the evaluation's first phase becomes a pipeline in target order -- the 4-lane Kahan state of
 * every (quiz, question, answer) is handed from shard to shard through the inboxes, tile of questions by tile, and the
 * last shard finishes the reference's own sum and publishes W_k to all shards. W_k is then bit-identical to a single
 * engine's (and CpuEngine's), so posteriors inside the evaluation are too and priorities meet the single-engine bar
 * (2e-12) instead of the summed-partials bar. Costs the pipeline fill: (shards-1)/tiles of phase 1. Same setting on every
 * shard. */
PQACORE_API void *PqaB200_P2PSetExactOrder(void *pvEngine, int32_t on);

/* ---- device-resident stepping and timing (bench.py "value" leg: no host<->device traffic inside) ---- */
/* Binds n quizzes as the resident batch: ids and one random draw per quiz are copied to the device once. */
PQACORE_API void *PqaB200_ResidentBind(void *pvEngine, int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms);
/* One NextQuestion pass (evaluation + selection) over the resident batch, asynchronous on the engine stream;
 * results stay on the device. */
PQACORE_API void *PqaB200_ResidentStep(void *pvEngine);
PQACORE_API void *PqaB200_ResidentFetch(void *pvEngine, int64_t *pQuestions /* n */);
/* Device time (ms, CUDA events on the engine stream) of the question-evaluation kernel inside the most recent
 * PqaB200_ResidentStep; waits for that kernel. -1 if no step was issued. */
PQACORE_API double PqaB200_ResidentLastEvalMs(void *pvEngine);
PQACORE_API void *PqaB200_Synchronize(void *pvEngine);
PQACORE_API void *PqaB200_EventCreate(void);
PQACORE_API void PqaB200_EventDestroy(void *pvEvent);
PQACORE_API void *PqaB200_EventRecord(void *pvEngine, void *pvEvent); /* on the engine's stream */
PQACORE_API void *PqaB200_EventSynchronize(void *pvEvent);
PQACORE_API double PqaB200_EventElapsedMs(void *pvStart, void *pvStop);
/* Number of kernel launches issued by this engine since creation (bench.py "gpu_launches"). */
PQACORE_API uint64_t PqaB200_KernelLaunchCount(void *pvEngine);
/* Writes a buffer larger than L2 on the engine's stream (timing hygiene between iterations). */
PQACORE_API void *PqaB200_FlushL2(void *pvEngine);

#ifdef __cplusplus
}
#endif
#endif /* PQA_B200_EXT_H */
