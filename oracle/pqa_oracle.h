/* TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.
 *
 * CPU oracle: a plain-C, scalar restatement of the srogatch/ProbQA CpuEngine<SRDoubleNumber> hot path
 * (StartQuiz, RecordAnswer, NextQuestion evaluation + selection, ListTopTargets, RecordQuizTarget/Train).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (probqa_b200/csrc) never links, loads or calls anything in oracle/.
 *
 * Parity pin: the primitives (Log2Hot, the 4-lane Kahan accumulator, PreciseSum/PairSum, SRHeapHelper::Down,
 * libstdc++ make_heap/pop_heap) and the whole-question evaluation are checked bit-for-bit against the
 * reference's own sources compiled by oracle/build_ref.sh into oracle/_ref/ (tests/test_oracle_vs_ref.py)
 * and against the reference's known-answer tests (SRAccumulatorTest.cpp:20-34, SRVectMathTest.cpp:45-104).
 *
 * All citations are relative to /root/reference/ProbQA/.
 *
 * Data layout (the reference's file layout, CpuEngine.cpp:664-688): sA flat [(i*K + k)*ldT + j],
 * mD flat [i*ldT + j], vB/prior [j]; ldT >= T is the row stride in doubles. Gap bitmaps are byte arrays,
 * bit x lives at byte x>>3, bit x&7 (SRBitArray); NULL means "no gaps". Lanes j >= T of the last 4-wide
 * vector behave as gaps (GapTracker.h:12-15 padding bits read as gap). */
#ifndef PQA_ORACLE_H
#define PQA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int64_t iTarget; double prob; } OraRatedTarget;     /* Interface/PqaCommon.h:54-61 */
typedef struct { int64_t iQuestion; int64_t iAnswer; } OraAnsweredQuestion; /* Interface/PqaCommon.h:48-52 */

/* 4 independent Kahan lanes: SRPlatform/Interface/SRAccumVectDbl256.h:18-46 */
typedef struct { double sum[4]; double corr[4]; } OraV4;
void   ora_v4_reset(OraV4 *a);
void   ora_v4_add(OraV4 *a, const double v[4]);                      /* SRAccumVectDbl256.h:40-46 */
void   ora_v4_add_at(OraV4 *a, int at, double v);                    /* SRAccumVectDbl256.h:48-54 */
double ora_v4_precise_sum(const OraV4 *a);                           /* SRAccumVectDbl256.h:62-92 */
double ora_v4_pair_sum(const OraV4 *a, const OraV4 *fellow, double *fellowSum); /* :94-133 */
double ora_v4_full_sum(const OraV4 *a);                              /* :56-60 */

/* SRPlatform/Interface/SRVectMath.h:87-135 with the table of SRPlatform/SRVectMath.cpp:30-44 */
double ora_log2hot(double x);
const double *ora_log2hot_table(void); /* 1024 entries */

/* SRPlatform/Interface/SRPoolRunner.h:96-110. Writes piece upper bounds, returns the number of pieces. */
int64_t ora_calc_split(int64_t nItems, int64_t nWorkers, int64_t *bounds);

/* CECreateQuizOperation.cpp:22-53 + CESetPriorsSubtaskSum.cpp:17-40 + Summator.h:11-21 +
 * CEDivTargPriorsSubtask.h:12-23.  W = worker count (reference: hardware_concurrency). */
void ora_start_quiz(const double *vB, const uint8_t *tgaps, int64_t T, int64_t W, double *prior);

/* CEQuiz.h:77-122 + CERecordAnswerSubtaskMul.cpp:15-42 (W here = the "loose" count max(1,hwc-1)). */
void ora_record_answer(const double *sArow, const double *mDrow, const uint8_t *tgaps, int64_t T, int64_t W,
                       double *prior);

/* ResumeQuiz: CECreateQuizOperation.cpp:55-83 + CEUpdatePriorsSubtaskMul.cpp:12-111 (mantissa/exponent-split product of
 * the answered questions' likelihoods; NOTE the reference multiplies by vB[j % 4] -- it loads the first vector of vB for
 * every target vector, :53 -- which is reproduced) + CpuEngine::NormalizePriors CpuEngine.cpp:284-335 with
 * CENormPriorsSubtaskMax.cpp / CENormPriorsSubtaskCorrSum.cpp + CEDivTargPriorsSubtask.h. W = worker count.
 * Returns 0, or 1 for the reference's I64Underflow error. nAQs must be >= 1 (0 is StartQuiz, BaseEngine.cpp:392-394). */
int ora_resume_quiz(const double *sA, const double *mD, const double *vB, int64_t ldT, int64_t K, int64_t T, int64_t W,
                    const uint8_t *tgaps, const OraAnsweredQuestion *aqs, int64_t nAQs, double *prior);

/* CEEvalQsSubtaskConsider.cpp:41-217 for ONE question (must be neither asked nor gap).
 * Outputs (any may be NULL): Wk/Hk/Vk [K], *lack, *totW. Returns the priority. */
double ora_eval_question(const double *sAi /* K rows, stride ldT */, const double *mDi, int64_t ldT,
                         const double *prior, const uint8_t *tgaps, int64_t K, int64_t T, int64_t nValidTargets,
                         double *Wk, double *Hk, double *Vk, double *lack, double *totW);

/* CpuEngine.cpp:337-374: all chunks of split(Q, 8*W); fills runLength[Q] (chunk-local Kahan prefixes),
 * priority[Q] (NaN for asked/gap questions; may be NULL), grand[<=8W] (Kahan prefix of chunk totals).
 * Returns the number of chunks. nThreads>1 runs chunks on pthreads (same results: chunks are independent). */
int64_t ora_eval_questions(const double *sA, const double *mD, int64_t ldT, const double *prior,
                           const uint8_t *asked, const uint8_t *qgaps, const uint8_t *tgaps,
                           int64_t Q, int64_t K, int64_t T, int64_t W, int nThreads,
                           double *runLength, double *priority, double *grand, int64_t *bounds);

/* CpuEngine.cpp:376-410 + BaseEngine.cpp:60-124, with the random 64-bit draw injected (SRDoubleNumber.h:35-39). */
int64_t ora_select_question(const double *runLength, const double *grand, const int64_t *bounds, int64_t nChunks,
                            int64_t Q, uint64_t rnd, const uint8_t *asked, const uint8_t *qgaps);
int64_t ora_find_nearest_question(int64_t iMiddle, int64_t Q, const uint8_t *asked, const uint8_t *qgaps);

/* CpuEngine.cpp:417-440 + CEListTopTargetsAlgorithm.cpp:30-97 + CEHeapifyPriorsSubtaskMake.cpp:42-87
 * (heapify branch; libstdc++ make_heap/pop_heap restated in C). Returns number listed. */
int64_t ora_list_top_targets(const double *prior, const uint8_t *tgaps, int64_t T, int64_t W, int64_t maxCount,
                             OraRatedTarget *dest);
/* 1 if CpuEngine.cpp:424-432 would pick the radix branch (not restated). */
int ora_would_use_radix(int64_t T, int64_t W, int64_t maxCount);

/* CpuEngine.cpp:442-466 + CETrainOperation.cpp:15-83 + CETrainTaskNumSpec.h:24-32: the sequential
 * RecordQuizTarget update (pairs of answers). */
void ora_record_quiz_target(double *sA, double *mD, double *vB, int64_t ldT, int64_t K,
                            const OraAnsweredQuestion *aqs, int64_t nAQs, int64_t iTarget, double amount);
/* CpuEngine.cpp:102-183 with W=1 semantics made deterministic: one bucket per (q % W), LIFO chains in
 * arrival order 0..n-1 (CETrainSubtaskDistrib.h:47-51, CETrainSubtaskAdd.cpp:17-38). */
void ora_train(double *sA, double *mD, double *vB, int64_t ldT, int64_t K,
               const OraAnsweredQuestion *aqs, int64_t nAQs, int64_t iTarget, double amount, int64_t W);

/* heap primitives exposed for cross-checking against libstdc++ (oracle/_ref) */
void ora_make_heap(OraRatedTarget *first, int64_t len);
void ora_pop_heap(OraRatedTarget *first, int64_t len);

#ifdef __cplusplus
}
#endif
#endif
