// probqa_b200: host side of the B200 engine. The method set mirrors the reference's IPqaEngine
// (PqaCore/Interface/IPqaEngine.h:13-114) for the quiz path, with the validation, error codes and quiz-registry
// behaviour of BaseEngine (PqaCore/BaseEngine.cpp) and the arithmetic of CpuEngine<SRDoubleNumber> executed by the
// kernels of pqa_kernels.cu / pqa_eval_staged.cu. There is no CPU compute path: every operation on priors, sA, mD
// and vB runs on the device; the host only validates, keeps the quiz registry (ids, answered lists, active
// question) and moves ids/answers/results through pinned staging buffers.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <cstdio>
#include <ctime>

#include "../../include/PqaB200Ext.h"
#include "pqa_errors.h"
#include "pqa_kernels.cuh"

namespace pqa {

struct CudaFail : std::runtime_error {
  int code; const char *file; int line;
  CudaFail(int c, const char *what, const char *f, int l) : std::runtime_error(what), code(c), file(f), line(l) {}
};
#define PQA_CU(expr)                                                                 \
  do {                                                                               \
    cudaError_t e_ = (expr);                                                         \
    if (e_ != cudaSuccess) throw ::pqa::CudaFail((int)e_, #expr, __FILE__, __LINE__); \
  } while (0)

template <typename T> class DevBuf {  // grow-only device array
 public:
  ~DevBuf() { if (p_) cudaFree(p_); }
  T *get() const { return p_; }
  size_t size() const { return n_; }
  void ensure(size_t n, cudaStream_t st = nullptr, bool keep = false);
 private:
  T *p_ = nullptr;
  size_t n_ = 0;
};

template <typename T> class PinBuf {  // grow-only pinned host array
 public:
  ~PinBuf() { if (p_) cudaFreeHost(p_); }
  T *get() const { return p_; }
  void ensure(size_t n);
 private:
  T *p_ = nullptr;
  size_t n_ = 0;
};

// GapTracker<TPqaId> (PqaCore/GapTracker.h:11-70): removed ids, reused LIFO.
class GapSet {
 public:
  bool IsGap(int64_t at) const { return isGap_[(size_t)at] != 0; }
  void Release(int64_t at) { isGap_[(size_t)at] = 1; gaps_.push_back(at); }
  int64_t Acquire() {                       // :38-49
    if (gaps_.empty()) { isGap_.push_back(0); return (int64_t)isGap_.size() - 1; }
    const int64_t a = gaps_.back();
    gaps_.pop_back();
    isGap_[(size_t)a] = 0;
    return a;
  }
  void GrowTo(int64_t n) { isGap_.resize((size_t)n, 0); }
  void Compact(int64_t n) { isGap_.assign((size_t)n, 0); gaps_.clear(); }
  int64_t GetNGaps() const { return (int64_t)gaps_.size(); }
  const std::vector<int64_t> &Gaps() const { return gaps_; }
  int64_t Size() const { return (int64_t)isGap_.size(); }
 private:
  std::vector<int64_t> gaps_;
  std::vector<uint8_t> isGap_;
};

// PermanentIdManager (PqaCore/PermanentIdManager.cpp): compact (array index) <-> permanent (never reused) ids.
class PermIds {
 public:
  int64_t PermFromComp(int64_t comp) const;
  int64_t CompFromPerm(int64_t perm) const;
  bool RemoveComp(int64_t comp);
  bool RenewComp(int64_t comp);
  bool GrowTo(int64_t nComp);
  bool OnCompact(int64_t nNew, const int64_t *pOldIds);
  bool EnsurePermIdGreater(int64_t bound);
  bool RemapPermId(int64_t srcPerm, int64_t destPerm);
  bool Save(FILE *f, bool empty = false) const;   // {i64 nextPermId, i64 nComp, nComp x i64}
  bool Load(FILE *f);
 private:
  int64_t nextPerm_ = 0;
  std::vector<int64_t> comp2perm_;
  std::unordered_map<int64_t, int64_t> perm2comp_;
};

template <typename T> inline void DevBuf<T>::ensure(size_t n, cudaStream_t st, bool keep) {
  if (n <= n_) return;
  size_t cap = std::max(n, n_ * 2);
  T *p = nullptr;
  PQA_CU(cudaMalloc(&p, cap * sizeof(T)));
  if (p_) {
    if (keep) PQA_CU(cudaMemcpyAsync(p, p_, n_ * sizeof(T), cudaMemcpyDeviceToDevice, st));
    PQA_CU(cudaStreamSynchronize(st));
    cudaFree(p_);
  }
  p_ = p; n_ = cap;
}
template <typename T> inline void PinBuf<T>::ensure(size_t n) {
  if (n <= n_) return;
  size_t cap = std::max(n, n_ * 2);
  T *p = nullptr;
  PQA_CU(cudaMallocHost(&p, cap * sizeof(T)));
  if (p_) { std::memcpy(p, p_, n_ * sizeof(T)); cudaFreeHost(p_); }
  p_ = p; n_ = cap;
}


// Makes an engine's device current for the scope of a call and restores the caller's: several engines of one process
// may sit on different GPUs (ShardGroup, tests), and a client thread's current device is whatever it used last.
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  explicit DeviceScope(int device) {
    if (device < 0) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceScope() { if (switched) cudaSetDevice(prev); }
  DeviceScope(const DeviceScope &) = delete;
  DeviceScope &operator=(const DeviceScope &) = delete;
};

void PlanCompactQuestions(const GapSet &gaps, int64_t n, int64_t *oldIds);   // CpuEngine.cpp:594-608
void PlanCompactTargets(const GapSet &gaps, int64_t n, int64_t *oldIds);     // CpuEngine.cpp:616-641
std::string HostLogicSelfTest();                                            // pqa_maint.cu; "" = ok

struct HostQuiz {                      // BaseQuiz.h:13-36 (host part); priors and asked bits live on the device
  bool present = false;
  int64_t activeQuestion = -1;
  std::vector<CiAnsweredQuestion> answers;
  mutable time_t lastUsage = 0;        // BaseQuiz::_lastUsage (BaseQuiz.h:21,34): refreshed by every call that uses the quiz
};

// One pending one-quiz call (NextQuestion / RecordAnswer / ListTopTargets) waiting to be combined with the calls other
// client threads make at the same time.
struct CallSlot {
  int kind = 0;                 // 0 NextQuestion, 1 RecordAnswer, 2 ListTopTargets, 3 StartQuiz, 4 RecordQuizTarget, 5 ReleaseQuiz
  int64_t quiz = -1, arg = 0;   // arg: answer (kind 1), maxCount (kind 2) or target (kind 4)
  double amount = 1.0;          // kind 4
  CiRatedTarget *dest = nullptr;
  int64_t result = -1;
  PqaError *err = nullptr;
  int defers = 0;                    // rounds this NextQuestion call was held back to let the cheap calls catch up
  bool done = false, lead = false;   // lead: the previous leader handed the leadership to this (still pending) call
  std::condition_variable cv;        // signalled for this call only (with Engine::combineMu_)
};

class ShardGroup;
class Engine {
  friend class ShardGroup;   // a group reads its shards' device views (kb(), pool()) to wire kernels across them
 public:
  Engine(const CiEngineDefinition &def, const CiB200Options &opts);
  virtual ~Engine();

  // --- IPqaEngine mirror (one quiz per call) ---
  virtual PqaError *Train(int64_t nQuestions, const CiAnsweredQuestion *pAQs, int64_t iTarget, double amount);
  int64_t StartQuiz(PqaError **err);
  virtual int64_t ResumeQuiz(PqaError **err, int64_t nAnswered, const CiAnsweredQuestion *pAQs);
  int64_t NextQuestion(PqaError **err, int64_t iQuiz);
  PqaError *RecordAnswer(int64_t iQuiz, int64_t iAnswer);
  int64_t GetActiveQuestionId(PqaError **err, int64_t iQuiz);
  PqaError *SetActiveQuestion(int64_t iQuiz, int64_t iQuestion);
  int64_t ListTopTargets(PqaError **err, int64_t iQuiz, int64_t maxCount, CiRatedTarget *pDest);
  PqaError *RecordQuizTarget(int64_t iQuiz, int64_t iTarget, double amount);
  PqaError *ReleaseQuiz(int64_t iQuiz);
  uint64_t GetTotalQuestionsAsked() const { return nQuestionsAsked_.load(std::memory_order_relaxed); }
  CiEngineDimensions CopyDims() const { return CiEngineDimensions{K_, Q_, T_}; }
  virtual PqaError *CopyATargets(int64_t iQuestion, int64_t iAnswer, int64_t maxTargets, double *pFreqs);
  virtual PqaError *CopyDTargets(int64_t iQuestion, int64_t maxTargets, double *pFreqs);
  virtual PqaError *CopyBTargets(int64_t maxTargets, double *pFreqs);
  virtual PqaError *SaveKB(const char *filePath);
  // BaseEngine::Shutdown (BaseEngine.cpp:260-322): no new operations, optional KB save, quizzes and KB released. Every later
  // call through the C ABI answers ObjectShutDown (pqa_cabi.cu checks IsShutDown before it enters the engine).
  virtual PqaError *Shutdown(const char *saveFilePath);
  bool IsShutDown() const { return shutdown_.load(std::memory_order_acquire); }
  virtual PqaError *SaveKBShard(const char *filePath, bool writeFrame);   // sharded engines: every shard writes its cells into one file
  static Engine *LoadKB(const char *filePath, const CiB200Options &opts, PqaError **err);

  // --- maintenance mode (BaseEngine.cpp:640-779, CpuEngine.cpp:468-658) and id maps (BaseEngine.cpp:154-218) ---
  virtual PqaError *ClearOldQuizzes(int64_t maxCount, double maxAgeSec);   // BaseEngine.cpp:814-872
  virtual PqaError *StartMaintenance(bool forceQuizzes);
  virtual PqaError *FinishMaintenance();
  virtual PqaError *AddQsTs(int64_t nQuestions, CiAddQorTParam *pAqps, int64_t nTargets, CiAddQorTParam *pAtps);
  virtual PqaError *RemoveQuestions(int64_t nQuestions, const int64_t *pQIds);
  virtual PqaError *RemoveTargets(int64_t nTargets, const int64_t *pTIds);
  virtual PqaError *Compact(int64_t *pnQuestions, const int64_t **ppOldQuestions, int64_t *pnTargets, const int64_t **ppOldTargets);
  bool MapIds(int kind, bool permFromComp, int64_t count, int64_t *pIds);   // kind: 0 questions, 1 targets, 2 quizzes
  bool EnsurePermQuizGreater(int64_t bound);
  bool RemapQuizPermId(int64_t srcPermId, int64_t destPermId);

  // --- batches of concurrent quizzes ---
  virtual PqaError *StartQuizBatch(int64_t n, int64_t *pQuizIds);
  virtual PqaError *ResumeQuizBatch(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, int64_t *pQuizIds);
  PqaError *ResumeQuizBatchEx(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, int64_t *pQuizIds,
                              const ResumeSource *src, const PoolList *pools, std::vector<int> *status);
  virtual PqaError *NextQuestionBatch(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms, int64_t *pQuestions,
                              void **ppErrors);
  virtual PqaError *RecordAnswerBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers);
  virtual PqaError *SetActiveQuestionBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pQuestions);
  virtual PqaError *ListTopTargetsBatch(int64_t n, const int64_t *pQuizIds, int64_t maxCount, CiRatedTarget *pDest,
                                int64_t *pCounts);
  virtual PqaError *RecordQuizTargetBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pTargets, const double *pAmounts);
  virtual PqaError *ReleaseQuizBatch(int64_t n, const int64_t *pQuizIds);

  // --- KB transfer / inspection ---
  virtual PqaError *UploadKB(const double *sA, const double *mD, const double *vB);
  virtual PqaError *DownloadKB(double *sA, double *mD, double *vB);
  virtual PqaError *CopyQuizPriors(int64_t iQuiz, double *pPriors);
  virtual PqaError *SetQuizPriors(int64_t iQuiz, const double *pPriors);
  virtual PqaError *EvalQuestions(int64_t n, const int64_t *pQuizIds, double *pPriorities, double *pRunLength,
                          double *pGrandTotals, int64_t *pnChunks);
  virtual PqaError *EvalQuestionsDetailed(int64_t iQuiz, double *pW, double *pH, double *pV, double *pLack, double *pPriorities);
  virtual PqaError *EvalQuestionsDetailedBatch(int64_t n, const int64_t *pQuizIds, double *pW, double *pH, double *pV, double *pLack,
                                               double *pPriorities);
  virtual PqaError *SetEvalKernel(int32_t which, int64_t chunkTargets, int64_t quizzesPerCta, int32_t kahanLanesPerThread);

  // --- question-sharded operation (PqaB200Ext.h) ---
  virtual PqaError *ShardEval(int64_t n, const int64_t *pQuizIds);
  virtual PqaError *ShardSelect(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms, int64_t *pQuestions, void **ppErrors);
  virtual PqaError *ShardRecordAnswerBegin(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers);
  virtual PqaError *ShardRecordAnswerEnd(int64_t n, const int64_t *pQuizIds);
  virtual PqaError *ShardBuffer(int32_t which, void **ppDevice, int64_t *pCount);
  int64_t questionShardFirst() const { return qFirst_; }
  int64_t questionShardCount() const { return qLocal_; }

  // --- target-sharded operation (PqaB200Ext.h) ---
  virtual PqaError *TShardEvalW(int64_t n, const int64_t *pQuizIds);
  virtual PqaError *TShardEvalHVL(int64_t n, const int64_t *pQuizIds);
  virtual PqaError *TShardPriority(int64_t n, const int64_t *pQuizIds);
  int64_t targetShardFirst() const { return tFirst_; }
  int64_t targetShardCount() const { return tLocal_; }
  bool IsTargetSharded() const { return tLocal_ != T_; }
  bool IsSharded() const { return qLocal_ != Q_ || tLocal_ != T_; }
  // --- shard exchange over peer memory (PqaB200Ext.h "P2P" entry points); works for question and target shards ---
  virtual PqaError *P2PInit(int32_t rank, int32_t nRanks, int64_t maxQuizzes, void **ppBase, int64_t *pBytes);
  virtual PqaError *P2PExportHandle(uint8_t *pHandle64);
  virtual PqaError *P2POpenHandle(const uint8_t *pHandle64, void **ppPeerBase);
  virtual PqaError *P2PConnect(void *const *pBases);
  virtual PqaError *P2PNextQuestionBegin(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms);
  virtual PqaError *P2PNextQuestionEnd(int64_t n, const int64_t *pQuizIds, int64_t *pQuestions, void **ppErrors);
  virtual PqaError *P2PRecordAnswerBegin(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers);
  virtual PqaError *P2PRecordAnswerEnd();
  // Warn-only conditions of the reference's evaluation counted by the selection kernels (DeviceKB::anomalies); a growing
  // counter is also reported on stderr, as the reference's logger does.
  virtual PqaError *AnomalyCounts(uint64_t *pCounts3);
  virtual PqaError *P2PLastPhaseMs(double *pMs5);   // device time of the stages of the last target-sharded P2PNextQuestion
  virtual PqaError *P2PSetExactOrder(int32_t on);   // target shards: hand the Kahan lanes from shard to shard (W_k bit-exact)
  // closed-form synthetic KB of SURVEY.md 8d written on the device (this engine's shard of it)
  virtual PqaError *FillBinarySearchKB(double rounds);

  // --- device-resident stepping ---
  virtual PqaError *ResidentBind(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms);
  virtual PqaError *ResidentStep();
  virtual PqaError *ResidentFetch(int64_t *pQuestions);
  virtual double ResidentLastEvalMs();   // device time of the evaluation kernel of the most recent ResidentStep (-1 on error)
  virtual PqaError *Synchronize();
  virtual PqaError *FlushL2();
  cudaStream_t stream() const { return stream_; }
  int device() const { return device_; }
  int emulatedWorkers() const { return W_; }

 protected:
  // A shell without device state: the host-side quiz registry, id maps and the call combiner of an engine whose KB lives
  // in other engines (ShardGroup, pqa_group.h).
  struct ShellTag {};
  Engine(const CiEngineDefinition &def, const CiB200Options &opts, ShellTag);
  void Submit(CallSlot &slot);                           // flat combining of concurrent one-quiz calls
  void RunCombined(const std::vector<CallSlot *> &batch);
  PqaError *CheckQuiz(int64_t iQuiz) const;              // BaseEngine::UseQuiz, BaseEngine.cpp:399-419
  PqaError *ValidateRecordAnswer(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) const;
  PqaError *FinishNextQuestion(int64_t m, const std::vector<int64_t> &valid, const std::vector<int64_t> &where,
                               int64_t *pQuestions, void **ppErrors, PqaError *firstErr);
  PqaError *WriteKBFile(FILE *f, const char *filePath, bool frame);
  PqaError *WrongMode(const char *what) const;          // regular-only operation called in maintenance mode
  void SyncGapBits();                                    // host gap trackers -> device bitmaps read by the kernels
  void ResizeKB(int64_t newQ, int64_t newT);             // grows sA/mD/vB on the device, keeps the old cells
  void DropQuizPool();
  bool OwnsQuestion(int64_t q) const { return q >= qFirst_ && q < qFirst_ + qLocal_; }
  int64_t AssignQuizId();                                // BaseEngine::AssignQuiz + GapTracker::Acquire
  void EnsureQuizCapacity(int64_t nSlots);
  void EnsureBatchScratch(int64_t n, bool needTop, int64_t maxCount);
  void UploadIds(int64_t n, const int64_t *ids);
  uint64_t NextRandom();
  DeviceKB kb() const;
  DeviceKB kbQuiz() const;
  // kb() for an evaluation launch: first brings the derived KB (pqa_eval_staged.cu) up to date with sA / mD -- all of it
  // after an upload / fill / load / resize / gap change, the touched questions after a Train / RecordQuizTarget.
  DeviceKB kbEval();
  bool UseFewPath(int64_t n) const;       // the fused one-launch NextQuestion of 1 .. eval_few_max() quizzes applies
  void EnsureFewResources();
  void MarkKBChanged() { derAllDirty_ = true; derDirtyList_.clear(); }
  void MarkQuestionsChanged(const TrainOp *ops, int64_t nOps);
  QuizPool pool() const;
  PqaError *ApplyTrain(const TrainOp *ops, int64_t nOps, const std::vector<int64_t> &targets,
                       const std::vector<double> &amounts);
  static int64_t AppendQuizOps(TrainOp *dst, const CiAnsweredQuestion *aqs, int64_t n, int64_t iTarget, double amount);

  mutable std::mutex mu_;
  std::mutex combineMu_;
  std::vector<CallSlot *> combinePending_;
  std::condition_variable gatherCv_;   // the leader waits here (briefly) for the rest of a cohort of concurrent callers
  bool combineLeader_ = false, gathering_ = false;
  // PQA_B200_STATS=1: combining statistics printed to stderr when the engine is released
  uint64_t statBatches_ = 0, statCalls_[6] = {}, statKindLaunches_[6] = {};
  double statRunSec_ = 0, statGatherSec_ = 0;
  size_t expectedBatch_ = 1;           // recent batch size: how many concurrent callers a leader may expect to show up
  int device_ = 0, W_ = 1, smCount_ = 148;
  int64_t Q_ = 0, K_ = 0, T_ = 0, Tp_ = 0, askedWords_ = 0;
  int64_t qFirst_ = 0, qLocal_ = 0;   // question shard held by this engine (0, Q_ for a single-device engine)
  int64_t tFirst_ = 0, tLocal_ = 0, TpL_ = 0;   // target shard (0, T_ unless target-sharded); TpL_ = row stride of sA/mD
  int64_t shardPriorityCount_ = 0, shardPriorsCount_ = 0, shardWCount_ = 0, shardHVLCount_ = 0;
  DevBuf<double> dShardPriority_, dShardPriors_, dShardW_, dShardHVL_;
  // peer-memory exchange state: own inbox, every shard's inbox base, lockstep counters (identical on all shards)
  int p2pRank_ = -1, p2pRanks_ = 0;
  int64_t p2pCap_ = 0;
  char *p2pInbox_ = nullptr;
  size_t p2pBytes_ = 0, p2pOffW_ = 0, p2pOffHVL_ = 0, p2pOffRows_ = 0, p2pOffPri_ = 0, p2pOffState_ = 0;
  size_t p2pSzW_ = 0, p2pSzHVL_ = 0, p2pSzRows_ = 0, p2pSzPri_ = 0, p2pSzState_ = 0;   // bytes of one slot / one parity copy
  bool p2pExactOrder_ = false, p2pSameDevicePeer_ = false;
  DevBuf<unsigned> dTileCounters_;
  char *p2pPeer_[kMaxPeers] = {};
  bool p2pOpened_[kMaxPeers] = {};
  bool p2pConnected_ = false, p2pPending_ = false;
  uint64_t p2pEpoch_ = 0, p2pOps_ = 0;
  double *p2pLastPriority_ = nullptr;
  cudaEvent_t p2pEv_[6] = {};
  bool p2pPhasesValid_ = false;
  P2PFlags p2pFlags() const;
  PqaError *P2PCheckError();
  std::atomic<bool> shutdown_{false};
  void ReleaseDeviceState();                                       // frees the KB, the derived KB and the quiz pool (Shutdown)
  std::atomic<bool> maintenance_{false};                           // MaintenanceSwitch mode (MaintenanceSwitch.h): Regular / Maintenance
  GapSet qGaps_, tGaps_;                                 // removed questions / targets (BaseEngine _questionGaps, _targetGaps)
  PermIds pimQ_, pimT_, pimQuiz_;
  DevBuf<uint32_t> dQGapBits_, dTGapBits_;
  double initAmount_ = 0;
  uint32_t precMantissa_ = 0; uint16_t precExponent_ = 0;   // kept only to write them back into a KB file header
  cudaStream_t stream_ = nullptr;
  EvalConfig evalCfg_;

  double *dSA_ = nullptr, *dMD_ = nullptr, *dVB_ = nullptr, *dLog2Tbl_ = nullptr;
  // derived KB: allocated at the first evaluation, rebuilt lazily (kbEval)
  double *dDerR_ = nullptr, *dDerL_ = nullptr;
  size_t derCapR_ = 0, derCapL_ = 0;
  bool derAllDirty_ = true;
  std::vector<int64_t> derDirtyList_;      // local question indices touched since the last rebuild (unique)
  std::vector<uint8_t> derDirtyMark_;
  DevBuf<int64_t> dDerList_;
  PinBuf<int64_t> hDerList_;
  // quiz pool
  int64_t quizCap_ = 0;
  double *dPriors_ = nullptr, *dLogPriors_ = nullptr, *dNormS_ = nullptr;
  uint64_t *dAsked_ = nullptr;
  int64_t *dActive_ = nullptr;
  std::vector<HostQuiz> quizzes_;
  std::vector<int64_t> quizGaps_;   // released ids, reused LIFO (GapTracker.h:38-49)

  // per-call staging
  DevBuf<int64_t> dIds_, dAnswers_, dQuestions_, dCounts_, dGroupStart_, dTargets_;
  DevBuf<uint64_t> dRandoms_;
  DevBuf<double> dPriority_, dRunLength_, dGrand_, dDetail_, dAmounts_, dRowScratch_;
  DevBuf<CiRatedTarget> dTop_, dTopScratch_;
  DevBuf<TrainOp> dOps_;
  DevBuf<uint8_t> dSortScratch_;
  PinBuf<int64_t> hIds_, hAnswers_, hQuestions_, hCounts_;
  PinBuf<uint64_t> hRandoms_;
  PinBuf<CiRatedTarget> hTop_;
  PinBuf<double> hRow_;
  PinBuf<TrainOp> hOps_;

  // fused small-batch path: mapped host memory {int64 questions[kFewHostSlots]; uint64 seq}, ticket counters, call sequence number
  static constexpr int kFewHostSlots = 64;
  volatile void *hFew_ = nullptr;
  void *dFewHost_ = nullptr;
  DevBuf<unsigned> dFewTickets_;
  unsigned long long *dAnom_ = nullptr;      // DeviceKB::anomalies
  uint64_t *hAnom_ = nullptr;                // pinned copy read after the batch paths' synchronisation
  uint64_t anomSeen_[kAnomalyKinds] = {};
  uint64_t anomWarnings_ = 0;
  void QueueAnomalyCopy();
  void ReportAnomalies(const volatile uint64_t *now);
  uint64_t fewSeq_ = 0;

  // resident batch
  int64_t residentN_ = 0;
  DevBuf<int64_t> dResIds_, dResQuestions_;
  DevBuf<uint64_t> dResRandoms_;
  DevBuf<double> dResPriority_, dResRunLength_;
  cudaEvent_t evEvalStart_ = nullptr, evEvalStop_ = nullptr;
  void *flushBuf_ = nullptr;
  size_t flushBytes_ = 0;

  uint64_t rng_[2];
  std::atomic<uint64_t> nQuestionsAsked_{0};
};

} // namespace pqa
