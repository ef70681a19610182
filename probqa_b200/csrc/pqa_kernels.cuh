// probqa_b200: device-side data views and kernel launchers (sm_100a).
// Every launcher cites the reference CPU code whose results it reproduces (paths relative to
// /root/reference/ProbQA/). No launcher synchronises; all work is enqueued on the given stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqa {

// Knowledge base resident in HBM. Rows are padded to a multiple of 4 doubles (32 B) like the reference's
// SRFastArray rows (CpuEngine.decl.h:31-37): sA[(i*K + k)*Tp + j], mD[i*Tp + j], vB[j].
// Padding lanes hold sA = 0, mD = 1, vB = 0 so that they produce the +0.0 the reference's gap mask produces
// (GapTracker.h:12-15, SRSimd.h:253-256) without a branch.
struct DeviceKB {
  double *sA;
  double *mD;
  double *vB;
  // Derived KB streamed by the throughput evaluation kernels (pqa_eval_staged.cu), per local question i and 4-target
  // vector v: dR[(i*nV + v)*K*4 + k*4 + lane] = sA/mD, dL[(i*nV + v)*(K+1)*4 + k*4 + lane] = log2(sA/mD) for k < K and
  // 1/mD^2 for k = K (nV = Tp/4). nullptr for views that never evaluate (kbQuiz of a target shard, K > 8).
  double *dR = nullptr;
  double *dL = nullptr;
  const double *log2tbl;   // 1024-entry table of SRVectMath.cpp:30-44
  const uint32_t *tgaps;   // target gap bitmap (bit j of word j>>5) or nullptr
  const uint32_t *qgaps;   // question gap bitmap or nullptr
  int64_t Q, K, T, Tp;
  int64_t nValidTargets;   // T - #target gaps (CpuEngine.cpp:351)
  // Question shard held by this device: rows of questions qFirst .. qFirst+qCount-1 (sA/mD are indexed by i - qFirst).
  // A single-device engine has qFirst = 0, qCount = Q.
  int64_t qFirst, qCount;
  // Warn-only anomaly counters the selection maintains (the reference logs these conditions and carries on):
  // [0] priorities of available questions that are <= 0 or not finite (CEEvalQsSubtaskConsider.cpp:209-211),
  // [1] non-finite chunk totals (CpuEngine.cpp:368-371), [2] grand totals <= 0 (:375-377). nullptr = not counted.
  unsigned long long *anomalies = nullptr;
};
constexpr int kAnomalyKinds = 3;

constexpr int kMaxPeers = 8;   // shard engines of one box

// Per-quiz state resident in HBM, addressed by quiz slot.
struct QuizPool {
  double *priors;          // [slot][Tp]  normalised posterior (CEQuiz.decl.h _pPriorMants); padding = +0
  double *logPriors;       // [slot][Tp]  log2 of priors (derived; rewritten whenever priors are)
  uint64_t *asked;         // [slot][askedWords]  bit i = question i already answered (CEBaseQuiz _isQAsked)
  int64_t *active;         // [slot] active question or -1 (BaseQuiz.h _activeQuestion)
  double *normS = nullptr; // [slot] normaliser of the row being updated (long rows: two-kernel update, pqa_kernels.cu)
  int64_t askedWords;      // ceil(Q/64)
  int64_t Tp;
};

struct PeerBufs {           // where a result goes / comes from when several devices cooperate
  int n = 0;                // 1 = this device's own buffer only (the caller sums across shards);
                            // N = one buffer per shard: as output, p[r] is this shard's slot in shard r's inbox (peer
                            //     memory over NVLink); as input, p[r] is shard r's slot in this shard's inbox, summed in
                            //     shard order
  double *p[kMaxPeers] = {};
};

struct EvalDetail {        // optional per-answer outputs of the question evaluation (may all be nullptr)
  double *W, *H, *V;       // [n][Q][K]
  double *lack;            // [n][Q]
};

void launch_fill_kb(const DeviceKB &kb, double initSqr, double initMD, double init1, cudaStream_t st);
// pack flat (stride T) host-layout rows into padded device rows and back
void launch_pad_rows(double *dst, const double *src, int64_t nRows, int64_t T, int64_t Tp, double padValue, cudaStream_t st);
void launch_unpad_rows(double *dst, const double *src, int64_t nRows, int64_t T, int64_t Tp, cudaStream_t st);

// CECreateQuizStart::UpdateLikelihoods (CECreateQuizOperation.cpp:22-53): priors = vB / KahanSum_W(vB); asked = 0.
void launch_start_quiz(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, int W, cudaStream_t st);
// CEQuiz::RecordAnswer (CEQuiz.h:77-122): uses the quiz' active question, sets its asked bit, clears active.
// dAnswers[n]; W = loose worker count max(1, hwc-1).
void launch_record_answer(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                          const int64_t *dAnswers, int W, cudaStream_t st);
// ResumeQuiz (CECreateQuizOperation.cpp:55-83, CEUpdatePriorsSubtaskMul.cpp, CpuEngine::NormalizePriors): quiz b answers
// questions dAqQ[dAqStart[b] .. dAqStart[b+1]) with dAqA; dStatus[b] = 1 on the reference's I64Underflow.
void launch_resume_quiz(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, const int64_t *dAqStart,
                        const int64_t *dAqQ, const int64_t *dAqA, int W, int *dStatus, cudaStream_t st);
// The same over the shard engines of one process (ShardGroup): the cells of an answered question are read from whichever
// shard holds them (peer pointers), the finished row goes into every shard's replica of the quiz.
struct ResumeSource {
  int nShards = 0;
  int64_t K = 0;
  const double *sA[kMaxPeers] = {}, *mD[kMaxPeers] = {};
  int64_t qFirst[kMaxPeers] = {}, qCount[kMaxPeers] = {};   // questions held by shard r
  int64_t tFirst[kMaxPeers] = {}, TpL[kMaxPeers] = {};      // first target and padded column count (row stride) of shard r
};
struct PoolList {
  int n = 0;
  QuizPool p[kMaxPeers];
};
void launch_resume_quiz_multi(const ResumeSource &src, const PoolList &pools, const double *dVB, const uint32_t *dTGaps, int64_t T,
                              int64_t n, const int64_t *dSlots, const int64_t *dAqStart, const int64_t *dAqQ, const int64_t *dAqA,
                              int W, int *dStatus, cudaStream_t st);
// Renormalisation-free refresh of logPriors after priors were overwritten from the host.
void launch_refresh_log_priors(const QuizPool &qp, int64_t n, const int64_t *dSlots, cudaStream_t st);

// CEEvalQsSubtaskConsider::Run (CEEvalQsSubtaskConsider.cpp:41-217) for every (quiz, question):
// dPriority[n*Q] (NaN where asked/gap).
//   which 1 = exact : every rounding of the reference reproduced (4-lane Kahan order, Log2Hot, IEEE divides);
//                     W/H/V/lack bit-identical to CpuEngine, priority up to libm pow/exp2/log differences.
//   which 2 = staged: the throughput kernel (bulk-async/TMA staging of the sA/mD slab into shared memory, log-split
//                     entropy, warp-tree sums); results within the tolerance stated in DESIGN.md.
//   which 0 = auto (= 2).
// chunkTargets: 0 = auto; otherwise forces the number of targets staged per shared-memory chunk (tests).
struct EvalConfig {
  int which = 0;
  int smCount = 148;
  int64_t chunkTargets = 0;
  int64_t quizzesPerCta = 0;   // 0 = auto
  int kahanLanesPerThread = 0; // staged kernel: 4 = one thread per quiz, 1 = four threads per quiz, 0 = auto by batch
  // question-sharded engines exchanging over peer memory: every priority (and the NaN of an asked question) of this
  // device's questions is also stored at the same [b*Q + i] position of these buffers (the other shards' inboxes), from
  // the evaluation kernel's own epilogue. Staged kernels only.
  PeerBufs mirror;
};
void launch_eval_questions(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                           double *dPriority, const EvalDetail &det, const EvalConfig &cfg, cudaStream_t st);
// Derived KB (kb.dR / kb.dL): element counts for this view, and the builder -- all local questions (dList = nullptr) or
// the nList local question indices in dList (the questions a Train / RecordQuizTarget touched).
void derived_kb_doubles(const DeviceKB &kb, size_t *nR, size_t *nL);
void launch_build_derived(const DeviceKB &kb, const int64_t *dList, int64_t nList, cudaStream_t st);
// CpuEngine::NextQuestionSpec (CpuEngine.cpp:337-415) after the evaluation: chunk run-lengths, grand totals,
// weighted draw with dRandoms[n], nearest unasked question; writes dQuestions[n] and the quiz' active question.
// dRunLength[n*Q] / dGrand[n*nChunks] optional outputs. setActive=0 leaves the quiz untouched.
void launch_select_question(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                            const double *dPriority, const uint64_t *dRandoms, int W, double *dRunLength,
                            double *dGrand, int64_t *dQuestions, int setActive, cudaStream_t st);
int64_t select_chunk_count(int64_t Q, int W);
// One launch for a whole NextQuestion of 1 .. eval_few_max() quizzes on an un-sharded engine (the reference ABI's one quiz
// per call): evaluation from the derived KB, selection by the last CTA of each quiz, chosen questions written to mapped
// host memory hostQuestions[n] and then `seq` to *hostSeq (the caller polls it). slots / randoms are host arrays passed by
// value; dTickets: eval_few_ticket_count() zeroed counters owned by the engine. hostQuestions = nullptr: evaluation only (no
// question is chosen, no quiz state changes); dGrand: optional [n][nChunks] grand totals.
int eval_few_max();      // quizzes per launch
int eval_few_inline();   // up to this many, ids and draws travel in the launch parameters (slots / randoms host arrays)
int eval_few_ticket_count();
void launch_eval_few_select(const DeviceKB &kb, const QuizPool &qp, int n, const int64_t *slots, const uint64_t *randoms, int W,
                            double *dPriority, double *dRunLength, unsigned *dTickets, int64_t *hostQuestions,
                            uint64_t *hostSeq, uint64_t seq, double *dGrand, const int64_t *dSlots, const uint64_t *dRandoms,
                            cudaStream_t st);
// CEListTopTargetsAlgorithm::RunHeapifyBased (CEListTopTargetsAlgorithm.cpp:30-97). dScratch: n*T 16-byte items
// (used only when T items do not fit in shared memory). dDest: n*maxCount {int64 iTarget; double prob}.
void launch_list_top_targets(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, int W,
                             int64_t maxCount, void *dScratch, void *dDest, int64_t *dCounts, cudaStream_t st);

// CETrainOperation (CETrainOperation.cpp:15-83). The host splits every Perform2/Perform1 of the reference's
// sequential loop into per-cell operations and groups them by (question, target) cell group, keeping sequence
// order inside a group; one thread applies one group.
struct TrainOp {
  int64_t q, a0, a1;       // a1 = -1: ProcessOne on (q,a0); a1 == a0: the doubled step (4b, 4b^2);
                           // a1 != a0: same question, two answers (CETrainOperation.cpp:38-47, D += 2*addend0)
  int64_t target;
  double amount;
};
void launch_train_ops(const DeviceKB &kb, const TrainOp *dOps, const int64_t *dGroupStart, int64_t nGroups,
                      cudaStream_t st);
// vB[target] += amount, one thread per distinct target applying its amounts in sequence order.
void launch_add_vb(const DeviceKB &kb, const int64_t *dTargets, const double *dAmounts, const int64_t *dGroupStart,
                   int64_t nGroups, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
// Target-sharded engines (SURVEY.md 8e "Targets" row; BASELINE config 4): a device holds the columns
// [tFirst, tFirst + kbLocal.T) of every sA/mD row (kbLocal: T = local target count, Tp = local row stride; Q, K and
// nValidTargets describe the WHOLE KB), quiz state is replicated with full-length priors. The evaluation is two-phase
// with one exchange between the phases:
//   phase 1  partial W_k[b][i][k]  = sum over local targets of lik (4-lane Kahan in the reference's lane order)
//   -------  sum over shards (NCCL all-reduce, or partials pushed into every peer's inbox over NVLink) -------
//   phase 2  partial sum post*log2 post [K], sum (post-prior)^2 [K], lack sum [1] per (quiz, question)
//   -------  sum over shards -------
//   priority epilogue (CEEvalQsSubtaskConsider.cpp:134-207) on every shard, identical bits everywhere.
// W_k is needed before phase 2 because log2(lik/W_k) sits in the denominator of the lack term.
// phase 1: outW.p[*][(b*Q + i)*K + k]
// inState / outState (optional): exact-order pipeline -- the Kahan lanes (s, c) per (quiz, question, answer, lane),
// [((i*K + k)*4 + lane)*n + b]*2 doubles, continue from the previous shard's hand-over and go to the next shard instead
// of being summed; the shard without outState finishes the reference's sum and writes the complete W_k to outW.
// kbLocal.qFirst / qCount (with sA/mD pointing at row qFirst) select the questions of one pipeline tile.
// In-kernel control of the exact-order pipeline. Questions are grouped into tiles of tileQ consecutive questions. A CTA
// first waits until waitFlags[tile] >= epoch (the previous shard has handed over that tile; nullptr = do not wait); when
// the last CTA of a tile is done (tileCounters, self-resetting) it publishes epoch into signalFlags[r][tile] for
// r < nSignal (the next shard's wait flags; on the last shard the W_k-ready flags of every shard, which phase 2 waits on).
struct PipeCtl {
  const uint64_t *waitFlags = nullptr;
  uint64_t *signalFlags[kMaxPeers] = {};
  int nSignal = 0;
  unsigned *tileCounters = nullptr;
  uint64_t epoch = 0, timeoutNs = 0;
  uint64_t *errFlag = nullptr;
  int64_t tileQ = 1;
};
void launch_eval_tshard_w(const DeviceKB &kbLocal, const QuizPool &qp, int64_t tFirst, int64_t n, const int64_t *dSlots,
                          const PeerBufs &outW, const EvalConfig &cfg, cudaStream_t st, const double *inState = nullptr,
                          double *outState = nullptr, const PipeCtl *pipe = nullptr);
int64_t tshard_quiz_tiles(int64_t n);   // gridDim.y of the target-sharded kernels for a batch of n quizzes
// phase 2: W = sum_r inW.p[r]; outHVL.p[*][(b*Q + i)*(2K+1) + {H_0..H_{K-1}, V_0..V_{K-1}, L}]
void launch_eval_tshard_hvl(const DeviceKB &kbLocal, const QuizPool &qp, int64_t tFirst, int64_t n, const int64_t *dSlots,
                            const PeerBufs &inW, const PeerBufs &outHVL, const EvalConfig &cfg, cudaStream_t st,
                            const PipeCtl *pipe = nullptr);   // pipe: only waitFlags / epoch / tileQ are used
// epilogue: priority[b*Q + i] (NaN where asked / gap) from the summed W and H/V/L; det optional
void launch_tshard_priority(const DeviceKB &kbLocal, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                            const PeerBufs &inW, const PeerBufs &inHVL, double *dPriority, const EvalDetail &det,
                            cudaStream_t st);
// RecordAnswer on a target shard: out.p[*][b*qp.Tp + tFirst + jl] = prior * (sA[q][a][jl] / mD[q][jl]) for the local
// targets (CERecordAnswerSubtaskMul.cpp:27-37, divide first); the other columns of the row are not touched.
void launch_tshard_record_answer_partial(const DeviceKB &kbLocal, const QuizPool &qp, int64_t tFirst, int64_t n,
                                         const int64_t *dSlots, const int64_t *dAnswers, const PeerBufs &out,
                                         cudaStream_t st);
// ... and its second half on the complete row dRows[b*Tp + j]: asked bit, active = -1, the reference's Kahan
// normalisation (bit-exact, same as launch_record_answer). kbFull: T/Tp of the whole KB.
void launch_tshard_record_answer_finish(const DeviceKB &kbFull, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                                        const double *dRows, int W, cudaStream_t st);
// Closed-form synthetic KB of SURVEY.md 8d (probqa_b200/synth.py binary_search_kb, bit-identical) written straight into
// this device's shard: sA = (init + rounds*[k == ans(i,j)])^2, mD = sum_k sA, vB = init + rounds.
void launch_fill_binary_search_kb(const DeviceKB &kbLocal, int64_t tFirst, int64_t Tglobal, double init, double rounds,
                                  cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
// Exchange between shard engines over peer memory (NVLink P2P stores / same-device pointers), no host round trip.
// Every engine owns an "inbox" allocation whose base pointers are known to all engines (cudaIpc handles between
// processes). A kernel's epilogue stores what the other shards need straight into their inboxes; the barrier kernel
// below then makes those stores visible: thread r publishes `epoch` into shard r's flag word for this rank
// (st.release.sys after a system fence) and spins until shard r's word in the own inbox reaches `epoch`
// (ld.acquire.sys). Gives up after timeoutNs and raises *errFlag (the host turns it into an error) instead of hanging.
struct P2PFlags {
  int rank, nRanks;
  uint64_t *flags[kMaxPeers];   // flags[r] = shard r's flag array (kMaxPeers words, indexed by the signalling rank)
  uint64_t *errFlag;            // own inbox
};
void preload_exchange_kernels(int K);   // loads every kernel of the exchanged call sequences now (see pqa_kernels.cu)
void launch_p2p_barrier(const P2PFlags &f, uint64_t epoch, uint64_t timeoutNs, cudaStream_t st);
// point-to-point halves of the barrier, for pipelines: wait until flags[0 .. nFlags) >= value; publish value into
// *targets.p[r] for r < targets.n (the pointers are uint64_t flag words in peer inboxes)
void launch_p2p_wait(const uint64_t *flags, int nFlags, uint64_t value, uint64_t *errFlag, uint64_t timeoutNs, cudaStream_t st);
void launch_p2p_signal(const PeerBufs &targets, uint64_t value, cudaStream_t st);
// question-sharded RecordAnswer: rows of quizzes whose answered question (dQuestions[b]) this device owns go to every
// peer's inbox rows [b*Tp ..]; after the barrier the other quizzes' rows come out of the own inbox into the pool.
void launch_p2p_push_prior_rows(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                                const int64_t *dQuestions, const PeerBufs &out, cudaStream_t st);
void launch_p2p_pull_prior_rows(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                                const int64_t *dQuestions, const double *dRows, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
// Maintenance (CpuEngine::AddQsTsSpec / CompactSpec, CpuEngine.cpp:468-658): KB resize, initial amounts, compaction.
// dst[r*dstStride + c] = src[r*srcStride + c] for r < nRows, c < nCols
void launch_copy_rows(double *dst, int64_t dstStride, const double *src, int64_t srcStride, int64_t nRows, int64_t nCols,
                      cudaStream_t st);
struct FillRect {           // base[(row0 + r)*stride + col0 + c] = value for r < nRows, c < nCols
  double *base;
  int64_t stride, row0, nRows, col0, nCols;
  double value;
};
void launch_fill_rects(const FillRect *dRects, int64_t nRects, cudaStream_t st);
// dst[(i*rowsPer + k)*dstStride + j] = src[(oldRow[i]*rowsPer + k)*srcStride + oldCol[j]] for i < nNewRows, j < nNewCols;
// columns nNewCols .. dstStride-1 receive padValue
void launch_gather_kb(double *dst, int64_t dstStride, const double *src, int64_t srcStride, const int64_t *dOldRow,
                      int64_t nNewRows, int64_t rowsPer, const int64_t *dOldCol, int64_t nNewCols, double padValue,
                      cudaStream_t st);

// Device-grouped forms for large batches (pqa_train_sort.cu): operations / (target, amount) pairs arrive in sequence
// order, a stable radix sort by cell groups them on the device. Same results as the host-grouped launchers above.
size_t train_sort_scratch_bytes(int64_t n);
void launch_train_ops_device_grouped(const DeviceKB &kb, const TrainOp *dOps, int64_t nOps, void *dScratch,
                                     size_t scratchBytes, cudaStream_t st);
void launch_add_vb_device_grouped(const DeviceKB &kb, const int64_t *dTargets, const double *dAmounts, int64_t n,
                                  void *dScratch, size_t scratchBytes, cudaStream_t st);

void launch_gather_prior_rows(const QuizPool &qp, int64_t n, const int64_t *dSlots, double *dBuf, cudaStream_t st);
void launch_scatter_prior_rows(const QuizPool &qp, int64_t n, const int64_t *dSlots, const double *dBuf, cudaStream_t st);
void launch_set_active(const QuizPool &qp, int64_t n, const int64_t *dSlots, const int64_t *dQuestions, cudaStream_t st);
void launch_flush_l2(void *buf, size_t bytes, cudaStream_t st);

uint64_t kernel_launch_count();

} // namespace pqa
