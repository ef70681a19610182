// probqa_b200: host side of the B200 engine (see pqa_engine.h). Citations are relative to /root/reference/ProbQA/.
#include "pqa_engine.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <random>
#include <thread>

namespace pqa {

#define SRC_LINE_STR2(x) #x
#define SRC_LINE_STR(x) SRC_LINE_STR2(x)
#define PQA_FILE_LINE "pqa_engine.cu(" SRC_LINE_STR(__LINE__) "): "

static int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// The 1024-entry Log2Hot table of SRPlatform/SRVectMath.cpp:30-44: log2 of each bucket's mid-point, entry 0 scaled
// so that Log2Hot(1) <= 0. std::log2 here is the same libm call the reference makes at start-up.
static void build_log2_table(double *tbl) {
  for (uint32_t i = 0; i < 1024; i++) {
    const uint64_t bitsMid = 0x3FF0000000000000ull | ((uint64_t)i << 42) | (1ull << 41);
    double mid;
    std::memcpy(&mid, &bitsMid, 8);
    tbl[i] = std::log2(mid);
  }
  tbl[0] *= 9.9999999999999927e-01;
}

Engine::Engine(const CiEngineDefinition &def, const CiB200Options &opts) {
  Q_ = def._nQuestions; K_ = def._nAnswers; T_ = def._nTargets;
  Tp_ = (T_ + 3) & ~3ll;
  askedWords_ = (Q_ + 63) >> 6;
  initAmount_ = def._initAmount;
  precMantissa_ = def._precMantissa; precExponent_ = def._precExponent;
  qFirst_ = 0; qLocal_ = Q_;
  if (opts._questionShardCount > 0) {
    if (opts._questionShardFirst < 0 || opts._questionShardFirst + opts._questionShardCount > Q_)
      throw std::runtime_error("probqa_b200: question shard [" + std::to_string(opts._questionShardFirst) + ", +" +
                               std::to_string(opts._questionShardCount) + ") is outside 0.." + std::to_string(Q_));
    qFirst_ = opts._questionShardFirst; qLocal_ = opts._questionShardCount;
  }
  tFirst_ = 0; tLocal_ = T_;
  if (opts._targetShardCount > 0) {
    if (qLocal_ != Q_) throw std::runtime_error("probqa_b200: an engine is sharded over questions or over targets, not both");
    if (opts._targetShardFirst < 0 || opts._targetShardFirst + opts._targetShardCount > T_ || (opts._targetShardFirst & 3) ||
        ((opts._targetShardCount & 3) && opts._targetShardFirst + opts._targetShardCount != T_))
      throw std::runtime_error("probqa_b200: target shard [" + std::to_string(opts._targetShardFirst) + ", +" +
                               std::to_string(opts._targetShardCount) + ") must lie in 0.." + std::to_string(T_) +
                               " and start/end on multiples of 4 targets (the last shard ends at T)");
    tFirst_ = opts._targetShardFirst; tLocal_ = opts._targetShardCount;
  }
  TpL_ = (tLocal_ + 3) & ~3ll;
  qGaps_.GrowTo(Q_); tGaps_.GrowTo(T_);     // BaseEngine::AfterStatisticsInit (BaseEngine.cpp:32-34)
  pimQ_.GrowTo(Q_); pimT_.GrowTo(T_);
  int nDev = 0;
  PQA_CU(cudaGetDeviceCount(&nDev));
  if (nDev <= 0) throw std::runtime_error("probqa_b200: no CUDA device is visible; this engine has no CPU path");
  device_ = opts._device;
  if (device_ < 0) device_ = env_int("PQA_B200_DEVICE", -1);
  if (device_ < 0) PQA_CU(cudaGetDevice(&device_));
  PQA_CU(cudaSetDevice(device_));
  cudaDeviceProp prop;
  PQA_CU(cudaGetDeviceProperties(&prop, device_));
  if (prop.major < 10)
    throw std::runtime_error("probqa_b200: device compute capability " + std::to_string(prop.major) + "." +
                             std::to_string(prop.minor) + " < 10.0; the kernels are built for sm_100a only");
  smCount_ = prop.multiProcessorCount;
  evalCfg_.smCount = smCount_;
  W_ = opts._emulatedWorkers;
  if (W_ <= 0) W_ = env_int("PQA_B200_EMULATED_WORKERS", 0);
  if (W_ <= 0) W_ = (int)std::thread::hardware_concurrency();  // BaseCpuEngine.cpp:21-22
  if (W_ <= 0) W_ = 1;
  uint64_t seed = opts._rngSeed;
  if (seed == 0) { std::random_device rd; seed = ((uint64_t)rd() << 32) ^ rd(); }
  rng_[0] = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  rng_[1] = (seed ^ 0xBF58476D1CE4E5B9ull) * 0x94D049BB133111EBull + 1;
  if (!rng_[0] && !rng_[1]) rng_[1] = 1;

  PQA_CU(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  PQA_CU(cudaMalloc(&dSA_, sizeof(double) * (size_t)(qLocal_ * K_ * TpL_)));
  PQA_CU(cudaMalloc(&dMD_, sizeof(double) * (size_t)(qLocal_ * TpL_)));
  PQA_CU(cudaMalloc(&dVB_, sizeof(double) * (size_t)Tp_));
  PQA_CU(cudaMalloc(&dLog2Tbl_, sizeof(double) * 1024));
  double tbl[1024];
  build_log2_table(tbl);
  PQA_CU(cudaMemcpyAsync(dLog2Tbl_, tbl, sizeof(tbl), cudaMemcpyHostToDevice, stream_));
  // CpuEngine.cpp:44-45,53,66,71,83: sA = init^2, mD = K*init^2, vB = init
  const double initSqr = initAmount_ * initAmount_;
  launch_fill_kb(kb(), initSqr, initSqr * (double)K_, initAmount_, stream_);
  if (IsTargetSharded()) launch_fill_kb(kbQuiz(), initSqr, initSqr * (double)K_, initAmount_, stream_);   // full-length vB
  EnsureQuizCapacity(opts._initialQuizCapacity > 0 ? opts._initialQuizCapacity : 256);
  PQA_CU(cudaMalloc(&dAnom_, sizeof(unsigned long long) * kAnomalyKinds));
  PQA_CU(cudaMemsetAsync(dAnom_, 0, sizeof(unsigned long long) * kAnomalyKinds, stream_));
  PQA_CU(cudaHostAlloc((void **)&hAnom_, sizeof(uint64_t) * kAnomalyKinds, cudaHostAllocDefault));
  std::memset(hAnom_, 0, sizeof(uint64_t) * kAnomalyKinds);
  if (K_ <= 8) {      // the derived KB of the throughput kernels (kbEval): allocated now, built at the first evaluation
    size_t nR = 0, nL = 0;
    derived_kb_doubles(kb(), &nR, &nL);
    PQA_CU(cudaMalloc(&dDerR_, sizeof(double) * (nR + nL)));     // one range: R, then L
    dDerL_ = dDerR_ + nR;
    derCapR_ = nR; derCapL_ = nL;
    dDerList_.ensure((size_t)qLocal_, stream_); hDerList_.ensure((size_t)qLocal_);
  }
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
}

Engine::Engine(const CiEngineDefinition &def, const CiB200Options &opts, ShellTag) {
  Q_ = def._nQuestions; K_ = def._nAnswers; T_ = def._nTargets;
  Tp_ = (T_ + 3) & ~3ll; TpL_ = Tp_;
  askedWords_ = (Q_ + 63) >> 6;
  initAmount_ = def._initAmount;
  precMantissa_ = def._precMantissa; precExponent_ = def._precExponent;
  qFirst_ = 0; qLocal_ = Q_; tFirst_ = 0; tLocal_ = T_;
  qGaps_.GrowTo(Q_); tGaps_.GrowTo(T_);
  pimQ_.GrowTo(Q_); pimT_.GrowTo(T_);
  device_ = opts._device;
  W_ = opts._emulatedWorkers;
  if (W_ <= 0) W_ = env_int("PQA_B200_EMULATED_WORKERS", 0);
  if (W_ <= 0) W_ = (int)std::thread::hardware_concurrency();
  if (W_ <= 0) W_ = 1;
  uint64_t seed = opts._rngSeed;
  if (seed == 0) { std::random_device rd; seed = ((uint64_t)rd() << 32) ^ rd(); }
  rng_[0] = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
  rng_[1] = (seed ^ 0xBF58476D1CE4E5B9ull) * 0x94D049BB133111EBull + 1;
  if (!rng_[0] && !rng_[1]) rng_[1] = 1;
}

Engine::~Engine() {
  DeviceScope devScope(stream_ ? device_ : -1);
  if (stream_) cudaStreamSynchronize(stream_);
  if (env_int("PQA_B200_STATS", 0) && statBatches_ > 0) {
    static const char *names[6] = {"NextQuestion", "RecordAnswer", "ListTopTargets", "StartQuiz", "RecordQuizTarget", "ReleaseQuiz"};
    fprintf(stderr, "[probqa_b200] combined batches: %llu, in-batch time %.3f s, gather-window time %.3f s\n",
            (unsigned long long)statBatches_, statRunSec_, statGatherSec_);
    for (int k = 0; k < 6; k++)
      if (statCalls_[k])
        fprintf(stderr, "[probqa_b200]   %-16s %10llu calls in %8llu launches (%.1f per launch)\n", names[k],
                (unsigned long long)statCalls_[k], (unsigned long long)statKindLaunches_[k],
                (double)statCalls_[k] / (double)statKindLaunches_[k]);
  }
  cudaFree(dSA_); cudaFree(dMD_); cudaFree(dVB_); cudaFree(dLog2Tbl_);
  cudaFree(dDerR_);
  if (hFew_) cudaFreeHost((void *)hFew_);
  if (hAnom_) cudaFreeHost(hAnom_);
  if (dAnom_) cudaFree(dAnom_);
  cudaFree(dPriors_); cudaFree(dLogPriors_); cudaFree(dAsked_); cudaFree(dActive_); cudaFree(dNormS_);
  for (int r = 0; r < kMaxPeers; r++) if (p2pOpened_[r]) cudaIpcCloseMemHandle(p2pPeer_[r]);
  if (p2pInbox_) cudaFree(p2pInbox_);
  if (flushBuf_) cudaFree(flushBuf_);
  if (evEvalStart_) { cudaEventDestroy(evEvalStart_); cudaEventDestroy(evEvalStop_); }
  for (cudaEvent_t e : p2pEv_) if (e) cudaEventDestroy(e);
  if (stream_) cudaStreamDestroy(stream_);
}

DeviceKB Engine::kb() const {
  DeviceKB k;
  k.sA = dSA_; k.mD = dMD_; k.vB = dVB_; k.log2tbl = dLog2Tbl_;
  k.dR = dDerR_; k.dL = dDerL_;
  // gap bitmaps exist only while something is removed (RemoveTargets / RemoveQuestions without a Compact yet)
  k.tgaps = tGaps_.GetNGaps() > 0 ? dTGapBits_.get() : nullptr;
  k.qgaps = qGaps_.GetNGaps() > 0 ? dQGapBits_.get() : nullptr;
  k.Q = Q_; k.K = K_; k.T = tLocal_; k.Tp = TpL_;   // a target-sharded engine sees its own columns here
  k.nValidTargets = T_ - tGaps_.GetNGaps();  // CpuEngine.cpp:351
  k.qFirst = qFirst_; k.qCount = qLocal_;
  k.anomalies = dAnom_;
  return k;
}
// The view the quiz-level kernels use (StartQuiz, selection, ListTopTargets, vB updates): whole-length vB / priors.
// Identical to kb() unless the engine is target-sharded, in which case it carries no sA/mD rows at all.
DeviceKB Engine::kbQuiz() const {
  DeviceKB k = kb();
  if (IsTargetSharded()) { k.sA = nullptr; k.mD = nullptr; k.dR = nullptr; k.dL = nullptr; k.qCount = 0; k.T = T_; k.Tp = Tp_; }
  return k;
}
// caller holds mu_ and has this engine's device current
DeviceKB Engine::kbEval() {
  if (K_ > 8) return kb();      // served by the exact kernel, which reads sA / mD
  size_t nR = 0, nL = 0;
  derived_kb_doubles(kb(), &nR, &nL);
  if (nR != derCapR_ || nL != derCapL_ || !dDerR_) {
    // Normally allocated by the constructor: allocations and frees synchronise the device, which must not happen while
    // another shard engine of this process spins in an exchange barrier on the same GPU. Only a KB that grew in maintenance
    // mode (single engines only) comes through here.
    PQA_CU(cudaStreamSynchronize(stream_));
    if (dDerR_) cudaFree(dDerR_);
    dDerR_ = dDerL_ = nullptr; derCapR_ = derCapL_ = 0;
    PQA_CU(cudaMalloc(&dDerR_, sizeof(double) * (nR + nL)));
    dDerL_ = dDerR_ + nR;
    derCapR_ = nR; derCapL_ = nL;

    dDerList_.ensure((size_t)qLocal_, stream_); hDerList_.ensure((size_t)qLocal_);
    derAllDirty_ = true;
  }
  if (derAllDirty_) {
    launch_build_derived(kb(), nullptr, 0, stream_);
  } else if (!derDirtyList_.empty()) {
    const int64_t nd = (int64_t)derDirtyList_.size();
    dDerList_.ensure((size_t)nd, stream_); hDerList_.ensure((size_t)nd);      // both sized for qLocal_ at creation
    PQA_CU(cudaStreamSynchronize(stream_));      // an earlier rebuild may still be reading the pinned list
    std::memcpy(hDerList_.get(), derDirtyList_.data(), sizeof(int64_t) * (size_t)nd);
    PQA_CU(cudaMemcpyAsync(dDerList_.get(), hDerList_.get(), sizeof(int64_t) * (size_t)nd, cudaMemcpyHostToDevice, stream_));
    launch_build_derived(kb(), dDerList_.get(), nd, stream_);
  }
  for (int64_t q : derDirtyList_) derDirtyMark_[(size_t)q] = 0;
  derDirtyList_.clear();
  derAllDirty_ = false;
  return kb();
}
void Engine::MarkQuestionsChanged(const TrainOp *ops, int64_t nOps) {
  if (derAllDirty_) return;
  if ((int64_t)derDirtyMark_.size() != qLocal_) derDirtyMark_.assign((size_t)qLocal_, 0);
  for (int64_t x = 0; x < nOps; x++) {
    const int64_t ql = ops[x].q - qFirst_;
    if (ql < 0 || ql >= qLocal_ || derDirtyMark_[(size_t)ql]) continue;
    derDirtyMark_[(size_t)ql] = 1;
    derDirtyList_.push_back(ql);
  }
  if ((int64_t)derDirtyList_.size() * 2 > qLocal_) {      // most questions touched: rebuild everything in one sweep
    for (int64_t q : derDirtyList_) derDirtyMark_[(size_t)q] = 0;
    MarkKBChanged();
  }
}
QuizPool Engine::pool() const {
  QuizPool p;
  p.priors = dPriors_; p.logPriors = dLogPriors_; p.asked = dAsked_; p.active = dActive_; p.normS = dNormS_;
  p.askedWords = askedWords_; p.Tp = Tp_;
  return p;
}

void Engine::EnsureQuizCapacity(int64_t nSlots) {
  if (nSlots <= quizCap_) return;
  const int64_t cap = std::max<int64_t>(std::max<int64_t>(nSlots, quizCap_ * 2), 64);
  double *np = nullptr, *nl = nullptr, *ns = nullptr; uint64_t *na = nullptr; int64_t *nact = nullptr;
  // + one vector: the evaluation kernels prefetch the priors one 4-target vector ahead, also past a row's end
  PQA_CU(cudaMalloc(&np, sizeof(double) * (size_t)(cap * Tp_ + 4)));
  PQA_CU(cudaMalloc(&nl, sizeof(double) * (size_t)(cap * Tp_ + 4)));
  PQA_CU(cudaMemsetAsync(np + cap * Tp_, 0, sizeof(double) * 4, stream_));
  PQA_CU(cudaMemsetAsync(nl + cap * Tp_, 0, sizeof(double) * 4, stream_));
  PQA_CU(cudaMalloc(&na, sizeof(uint64_t) * (size_t)(cap * askedWords_)));
  PQA_CU(cudaMalloc(&nact, sizeof(int64_t) * (size_t)cap));
  PQA_CU(cudaMalloc(&ns, sizeof(double) * (size_t)cap));
  if (quizCap_ > 0) {
    PQA_CU(cudaMemcpyAsync(np, dPriors_, sizeof(double) * (size_t)(quizCap_ * Tp_), cudaMemcpyDeviceToDevice, stream_));
    PQA_CU(cudaMemcpyAsync(nl, dLogPriors_, sizeof(double) * (size_t)(quizCap_ * Tp_), cudaMemcpyDeviceToDevice, stream_));
    PQA_CU(cudaMemcpyAsync(na, dAsked_, sizeof(uint64_t) * (size_t)(quizCap_ * askedWords_), cudaMemcpyDeviceToDevice, stream_));
    PQA_CU(cudaMemcpyAsync(nact, dActive_, sizeof(int64_t) * (size_t)quizCap_, cudaMemcpyDeviceToDevice, stream_));
    PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
    cudaFree(dPriors_); cudaFree(dLogPriors_); cudaFree(dAsked_); cudaFree(dActive_); cudaFree(dNormS_);
  }
  dPriors_ = np; dLogPriors_ = nl; dAsked_ = na; dActive_ = nact; dNormS_ = ns;
  quizCap_ = cap;
}

uint64_t Engine::NextRandom() {  // xorshift128+, the generator family of SRFastRandom (SRFastRandom.h:31-42)
  uint64_t s1 = rng_[0];
  const uint64_t s0 = rng_[1];
  rng_[0] = s0;
  s1 ^= s1 << 23;
  rng_[1] = s1 ^ s0 ^ (s1 >> 18) ^ (s0 >> 5);
  return rng_[1] + s0;
}

PqaError *Engine::CheckQuiz(int64_t iQuiz) const {
  const int64_t nQuizzes = (int64_t)quizzes_.size();
  if (iQuiz < 0 || iQuiz >= nQuizzes)
    return ErrIndexOutOfRange(iQuiz, 0, nQuizzes - 1, PQA_FILE_LINE "Quiz index is not in quiz registry range.");
  if (!quizzes_[iQuiz].present)
    return ErrAbsentId(iQuiz, PQA_FILE_LINE "Quiz index is not in the registry (but rather at a gap).");
  quizzes_[iQuiz].lastUsage = std::time(nullptr);   // BaseQuiz::OnUsage, BaseEngine.cpp:416
  return nullptr;
}

int64_t Engine::AssignQuizId() {
  int64_t id;
  if (!quizGaps_.empty()) { id = quizGaps_.back(); quizGaps_.pop_back(); pimQuiz_.RenewComp(id); }   // BaseEngine.cpp:781-794
  else { id = (int64_t)quizzes_.size(); quizzes_.emplace_back(); pimQuiz_.GrowTo((int64_t)quizzes_.size()); }
  HostQuiz &q = quizzes_[id];
  q.present = true; q.activeQuestion = -1; q.answers.clear();
  q.lastUsage = std::time(nullptr);
  return id;
}

void Engine::UploadIds(int64_t n, const int64_t *ids) {
  hIds_.ensure(n); dIds_.ensure(n, stream_);
  std::memcpy(hIds_.get(), ids, sizeof(int64_t) * (size_t)n);
  PQA_CU(cudaMemcpyAsync(dIds_.get(), hIds_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
}

#define PQA_TRY try {
#define PQA_CATCH_RETURN_ERR                                                              \
  } catch (const CudaFail &cf) { return ErrCuda(cf.code, cf.what(), cf.file, cf.line);    \
  } catch (const std::exception &ex) { return ErrStd(ex.what()); }
#define PQA_CATCH_SET_ERR(ret)                                                            \
  } catch (const CudaFail &cf) { *err = ErrCuda(cf.code, cf.what(), cf.file, cf.line); return ret; \
  } catch (const std::exception &ex) { *err = ErrStd(ex.what()); return ret; }

// ---------------------------------------------------------------------------------------------------------
// StartQuiz: CpuEngine::StartQuiz -> CreateQuizInternal -> CECreateQuizStart::UpdateLikelihoods
// (CpuEngine.cpp:185-275, CECreateQuizOperation.cpp:22-53)
PqaError *Engine::StartQuizBatch(int64_t n, int64_t *pQuizIds) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds");
  if (maintenance_) return WrongMode("Start/Resume quiz");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return WrongMode("Start/Resume quiz");   // re-checked under the lock: StartMaintenance may have completed meanwhile
  PQA_TRY
  for (int64_t x = 0; x < n; x++) pQuizIds[x] = AssignQuizId();
  EnsureQuizCapacity((int64_t)quizzes_.size());
  UploadIds(n, pQuizIds);
  launch_start_quiz(kbQuiz(), pool(), n, dIds_.get(), W_, stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

int64_t Engine::StartQuiz(PqaError **err) {
  if (IsSharded()) { int64_t id = -1; *err = StartQuizBatch(1, &id); return *err ? -1 : id; }
  if (maintenance_) { *err = WrongMode("Start/Resume quiz"); return -1; }
  CallSlot s; s.kind = 3;
  Submit(s);                    // combined with the concurrent one-quiz calls of other client threads
  *err = s.err;
  return s.err ? -1 : s.result;
}

// ResumeQuiz: BaseEngine::ResumeQuiz (BaseEngine.cpp:386-397) -> CpuEngine::ResumeQuizSpec (CpuEngine.cpp:277-282) ->
// CreateQuizInternal (:185-270: index validation, asked bits, answers) -> CECreateQuizResume::UpdateLikelihoods.
// pCounts[x] answered questions of quiz x are taken from pAQs in order. A quiz with zero answers is a StartQuiz (:392-394).
PqaError *Engine::ResumeQuizBatch(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, int64_t *pQuizIds) {
  if (IsSharded()) return ErrNotImplemented("ResumeQuiz on a sharded engine (a ShardGroup resumes quizzes for its shards)");
  return ResumeQuizBatchEx(n, pCounts, pAQs, pQuizIds, nullptr, nullptr, nullptr);
}

// ResumeQuiz for one engine (src = nullptr), or as one member of a ShardGroup: the leader (src / pools given) runs the
// kernel once over the cells of all shards and stores the finished rows into every shard's quiz pool; a follower
// (followStatus given: the leader's per-resumed-quiz status) only keeps its registry in step and starts the quizzes that
// have no answers. outStatus (leader): receives that status array.
PqaError *Engine::ResumeQuizBatchEx(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, int64_t *pQuizIds,
                                    const ResumeSource *src, const PoolList *pools, std::vector<int> *status) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pCounts || !pQuizIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pCounts/pQuizIds");
  if (maintenance_) return WrongMode("Start/Resume quiz");
  const bool follower = src == nullptr && status != nullptr;
  int64_t total = 0;
  for (int64_t x = 0; x < n; x++) {
    if (pCounts[x] < 0) return ErrNegativeCount(pCounts[x], "|nAnswered| must be non-negative.");
    total += pCounts[x];
  }
  if (total > 0 && !pAQs) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pAQs");
  for (int64_t x = 0; x < total; x++) {   // CpuEngine.cpp:219-231 (reported there through an aggregate error)
    PqaError *inner = nullptr;
    if (pAQs[x]._iQuestion < 0 || pAQs[x]._iQuestion >= Q_)
      inner = ErrIndexOutOfRange(pAQs[x]._iQuestion, 0, Q_ - 1, PQA_FILE_LINE "Question index is not in KB range.");
    else if (pAQs[x]._iAnswer < 0 || pAQs[x]._iAnswer >= K_)
      inner = ErrIndexOutOfRange(pAQs[x]._iAnswer, 0, K_ - 1, PQA_FILE_LINE "Answer index is not in KB range.");
    if (inner) {
      PqaError *agg = MakeError(ErrCode::Aggregate, "", "Aggregate error [" + inner->ToString(true) + "]");
      delete inner;
      return agg;
    }
  }
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  for (int64_t x = 0; x < n; x++) pQuizIds[x] = AssignQuizId();
  EnsureQuizCapacity((int64_t)quizzes_.size());
  // quizzes without answers start; the others resume
  std::vector<int64_t> startIds, resumeIds, aqStart{0}, aqQ, aqA;
  int64_t off = 0;
  for (int64_t x = 0; x < n; x++) {
    if (pCounts[x] == 0) { startIds.push_back(pQuizIds[x]); continue; }
    resumeIds.push_back(pQuizIds[x]);
    HostQuiz &q = quizzes_[pQuizIds[x]];
    for (int64_t y = 0; y < pCounts[x]; y++) {
      aqQ.push_back(pAQs[off + y]._iQuestion); aqA.push_back(pAQs[off + y]._iAnswer);
      q.answers.push_back(pAQs[off + y]);                                     // CpuEngine.cpp:236-241
    }
    off += pCounts[x];
    aqStart.push_back((int64_t)aqQ.size());
  }
  if (!startIds.empty()) {
    UploadIds((int64_t)startIds.size(), startIds.data());
    launch_start_quiz(kbQuiz(), pool(), (int64_t)startIds.size(), dIds_.get(), W_, stream_);
    PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  }
  PqaError *result = nullptr;
  if (!resumeIds.empty()) {
    const int64_t m = (int64_t)resumeIds.size();
    std::vector<int> st((size_t)m, 0);
    if (follower) {
      if ((int64_t)status->size() != m) return MakeError(ErrCode::Internal, PQA_FILE_LINE "the shards' resume batches have diverged");
      st = *status;
    } else {
      UploadIds(m, resumeIds.data());
      dGroupStart_.ensure(aqStart.size(), stream_); dTargets_.ensure(aqQ.size(), stream_); dAnswers_.ensure(aqA.size(), stream_);
      dCounts_.ensure((size_t)m, stream_);   // reused as the int status array (m ints fit in m int64 slots)
      PQA_CU(cudaMemcpyAsync(dGroupStart_.get(), aqStart.data(), sizeof(int64_t) * aqStart.size(), cudaMemcpyHostToDevice, stream_));
      PQA_CU(cudaMemcpyAsync(dTargets_.get(), aqQ.data(), sizeof(int64_t) * aqQ.size(), cudaMemcpyHostToDevice, stream_));
      PQA_CU(cudaMemcpyAsync(dAnswers_.get(), aqA.data(), sizeof(int64_t) * aqA.size(), cudaMemcpyHostToDevice, stream_));
      if (src) {
        const DeviceKB kq = kbQuiz();
        launch_resume_quiz_multi(*src, *pools, kq.vB, kq.tgaps, T_, m, dIds_.get(), dGroupStart_.get(), dTargets_.get(),
                                 dAnswers_.get(), W_, reinterpret_cast<int *>(dCounts_.get()), stream_);
      } else {
        launch_resume_quiz(kb(), pool(), m, dIds_.get(), dGroupStart_.get(), dTargets_.get(), dAnswers_.get(), W_,
                           reinterpret_cast<int *>(dCounts_.get()), stream_);
      }
      PQA_CU(cudaMemcpyAsync(st.data(), dCounts_.get(), sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, stream_));
      PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
      if (status) *status = st;
    }
    for (int64_t x = 0; x < m; x++) {
      if (st[(size_t)x] == 0) continue;
      // CpuEngine.cpp:316-319 -> CreateQuizInternal unassigns the quiz (:257-261)
      HostQuiz &q = quizzes_[resumeIds[x]];
      q.present = false; q.answers.clear(); q.activeQuestion = -1;
      quizGaps_.push_back(resumeIds[x]);
      pimQuiz_.RemoveComp(resumeIds[x]);
      for (int64_t y = 0; y < n; y++) if (pQuizIds[y] == resumeIds[x]) pQuizIds[y] = -1;
      if (!result) result = MakeError(ErrCode::I64Underflow, "Max exponent over the priors is too low. Are all the targets in gaps?",
                                      "actual=<underflow>, minAllowed=<see CpuEngine.cpp:315>");
    }
  }
  return result;
  PQA_CATCH_RETURN_ERR
}

int64_t Engine::ResumeQuiz(PqaError **err, int64_t nAnswered, const CiAnsweredQuestion *pAQs) {
  if (nAnswered < 0) { *err = ErrNegativeCount(nAnswered, "|nAnswered| must be non-negative."); return -1; }
  int64_t id = -1;
  *err = ResumeQuizBatch(1, &nAnswered, pAQs, &id);
  return *err ? -1 : id;
}

// ---------------------------------------------------------------------------------------------------------
// NextQuestion: BaseEngine::NextQuestion (BaseEngine.cpp:421-437) -> CpuEngine::NextQuestionSpec (CpuEngine.cpp:337-415)
PqaError *Engine::NextQuestionBatch(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms, int64_t *pQuestions,
                                    void **ppErrors) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pQuestions) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pQuestions");
  if (IsSharded()) return ErrNotImplemented("NextQuestion on a sharded engine: use the PqaB200_Shard* protocol");
  if (maintenance_) return WrongMode("compute next question");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return WrongMode("compute next question");   // re-checked under the lock: StartMaintenance may have completed meanwhile
  PQA_TRY
  // validate; quizzes that fail validation get their own error and are left out of the launch
  std::vector<int64_t> valid; valid.reserve(n);
  std::vector<int64_t> where; where.reserve(n);
  PqaError *firstErr = nullptr;
  for (int64_t x = 0; x < n; x++) {
    pQuestions[x] = -1;
    if (ppErrors) ppErrors[x] = nullptr;
    PqaError *e = CheckQuiz(pQuizIds[x]);
    if (e) {
      if (ppErrors) ppErrors[x] = e;
      else if (!firstErr) firstErr = e;
      else delete e;
      continue;
    }
    valid.push_back(pQuizIds[x]); where.push_back(x);
  }
  const int64_t m = (int64_t)valid.size();
  if (m > 0 && UseFewPath(m)) {
    // The reference ABI's shape (one quiz per call, or the handful a combiner round collects): one fused launch, ids and
    // draws by value, results through mapped host memory (k_eval_few, pqa_eval_staged.cu).
    EnsureFewResources();
    hRandoms_.ensure(m);
    uint64_t *rnd = hRandoms_.get();
    for (int64_t x = 0; x < m; x++) rnd[x] = pRandoms ? pRandoms[where[x]] : NextRandom();
    dPriority_.ensure((size_t)(m * Q_), stream_); dRunLength_.ensure((size_t)(m * Q_), stream_);
    if (m > eval_few_inline()) {            // ids and draws of a larger batch go through device arrays
      UploadIds(m, valid.data());
      dRandoms_.ensure(m, stream_);
      PQA_CU(cudaMemcpyAsync(dRandoms_.get(), rnd, sizeof(uint64_t) * (size_t)m, cudaMemcpyHostToDevice, stream_));
    }
    const uint64_t seq = ++fewSeq_;
    launch_eval_few_select(kbEval(), pool(), (int)m, valid.data(), rnd, W_, dPriority_.get(), dRunLength_.get(),
                           dFewTickets_.get(), (int64_t *)dFewHost_, (uint64_t *)dFewHost_ + kFewHostSlots, seq, nullptr,
                           dIds_.get(), dRandoms_.get(), stream_);
    PQA_CU(cudaGetLastError());
    volatile uint64_t *seen = (volatile uint64_t *)hFew_ + kFewHostSlots;
    for (uint64_t spins = 0; *seen != seq; spins++) {
      if ((spins & 0xFFF) == 0xFFF) {          // every few microseconds: has the stream died or finished without an answer?
        const cudaError_t qe = cudaStreamQuery(stream_);
        if (qe == cudaSuccess) { if (*seen == seq) break; throw CudaFail((int)cudaErrorUnknown, "k_eval_few finished without publishing its result", __FILE__, __LINE__); }
        if (qe != cudaErrorNotReady) PQA_CU(qe);
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    ReportAnomalies((const volatile uint64_t *)hFew_ + kFewHostSlots + 1);
    uint64_t nAsked = 0;
    for (int64_t x = 0; x < m; x++) {
      const int64_t qst = ((volatile int64_t *)hFew_)[x];
      if (qst < 0) {  // CpuEngine.cpp:407-411
        PqaError *e = MakeError(ErrCode::QuestionsExhausted, PQA_FILE_LINE "Found no unasked question that is not in a gap.");
        if (ppErrors) ppErrors[where[x]] = e;
        else if (!firstErr) firstErr = e;
        else delete e;
        continue;
      }
      pQuestions[where[x]] = qst;
      quizzes_[valid[x]].activeQuestion = qst;  // :412
      nAsked++;                                  // :413
    }
    nQuestionsAsked_.fetch_add(nAsked, std::memory_order_relaxed);
  } else if (m > 0) {
    UploadIds(m, valid.data());
    hRandoms_.ensure(m); dRandoms_.ensure(m, stream_);
    for (int64_t x = 0; x < m; x++) hRandoms_.get()[x] = pRandoms ? pRandoms[where[x]] : NextRandom();
    PQA_CU(cudaMemcpyAsync(dRandoms_.get(), hRandoms_.get(), sizeof(uint64_t) * (size_t)m, cudaMemcpyHostToDevice, stream_));
    dPriority_.ensure((size_t)(m * Q_), stream_); dRunLength_.ensure((size_t)(m * Q_), stream_);
    dQuestions_.ensure(m, stream_); hQuestions_.ensure(m);
    EvalDetail det{nullptr, nullptr, nullptr, nullptr};
    launch_eval_questions(kbEval(), pool(), m, dIds_.get(), dPriority_.get(), det, evalCfg_, stream_);
    launch_select_question(kbQuiz(), pool(), m, dIds_.get(), dPriority_.get(), dRandoms_.get(), W_, dRunLength_.get(),
                           nullptr, dQuestions_.get(), 1, stream_);
    PQA_CU(cudaMemcpyAsync(hQuestions_.get(), dQuestions_.get(), sizeof(int64_t) * (size_t)m, cudaMemcpyDeviceToHost, stream_));
    QueueAnomalyCopy();
    PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
    ReportAnomalies(hAnom_);
    uint64_t nAsked = 0;
    for (int64_t x = 0; x < m; x++) {
      const int64_t qst = hQuestions_.get()[x];
      if (qst < 0) {  // CpuEngine.cpp:407-411
        PqaError *e = MakeError(ErrCode::QuestionsExhausted, PQA_FILE_LINE "Found no unasked question that is not in a gap.");
        if (ppErrors) ppErrors[where[x]] = e;
        else if (!firstErr) firstErr = e;
        else delete e;
        continue;
      }
      pQuestions[where[x]] = qst;
      quizzes_[valid[x]].activeQuestion = qst;  // :412
      nAsked++;                                  // :413
    }
    nQuestionsAsked_.fetch_add(nAsked, std::memory_order_relaxed);
  }
  return firstErr;
  PQA_CATCH_RETURN_ERR
}

// ---------------------------------------------------------------------------------------------------------
// Combining of concurrent one-quiz calls. The reference ABI is synchronous and one quiz per call, and its clients
// (PqaClient.cpp:238-245, a web front-end) call it from many threads at once. The thread that finds no batch in flight
// becomes the leader and executes everything that queued up -- its own call plus the calls other threads made while the
// previous batch was on the GPU -- as ONE batch launch per kind; then it hands leadership over. A single caller gets a
// batch of one with no added latency; under load the batch size grows by itself and the KB is streamed once per batch.
void Engine::Submit(CallSlot &slot) {
  std::unique_lock<std::mutex> lk(combineMu_);
  combinePending_.push_back(&slot);
  if (combineLeader_) {
    if (gathering_) gatherCv_.notify_one();
    // wait for this call's result, or for the leadership (every call has its own condition variable: finishing a batch
    // wakes exactly the callers it served)
    slot.cv.wait(lk, [&] { return slot.done || slot.lead; });
    if (slot.done) return;
  }
  combineLeader_ = true;                     // this call is in combinePending_, so it is part of a batch below
  while (!slot.done) {
    // Callers that were served by the previous batch come back with their next call within microseconds of each other.
    // When recent batches had company, give the cohort a short window to arrive instead of launching for the first one
    // alone (a lone caller never waits: expectedBatch_ stays 1).
    if (expectedBatch_ > 1 && combinePending_.size() < expectedBatch_) {
      gathering_ = true;
      const auto g0 = std::chrono::steady_clock::now();
      gatherCv_.wait_for(lk, std::chrono::microseconds(120), [&] { return combinePending_.size() >= expectedBatch_; });
      statGatherSec_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - g0).count();
      gathering_ = false;
    }
    std::vector<CallSlot *> batch;
    batch.swap(combinePending_);
    expectedBatch_ = std::max<size_t>(1, std::max(batch.size() - batch.size() / 8, expectedBatch_ - (expectedBatch_ + 3) / 4));
    // NextQuestion is the expensive call and its cost hardly depends on the batch size, while clients drift apart (a
    // client that finishes a quiz makes three cheap calls before its next NextQuestion). So when cheap calls are pending
    // too, the NextQuestion calls are held back for up to three rounds: the cheap calls run at once, their callers come
    // back with the next call, and the cohort meets again at one big NextQuestion launch.
    size_t nNext = 0, nOther = 0;
    int maxDefers = 0;
    for (CallSlot *c : batch) {
      if (c->kind == 0) { nNext++; maxDefers = std::max(maxDefers, c->defers); } else nOther++;
    }
    if (nNext > 0 && nOther > 0 && maxDefers < 3) {
      std::vector<CallSlot *> now;
      for (CallSlot *c : batch) {
        if (c->kind == 0) { c->defers++; combinePending_.push_back(c); } else now.push_back(c);
      }
      batch.swap(now);
    }
    lk.unlock();
    const auto r0 = std::chrono::steady_clock::now();
    try {
      RunCombined(batch);
    } catch (const std::exception &ex) {     // e.g. bad_alloc in the host bookkeeping: every caller of the batch hears about it
      for (CallSlot *c : batch) if (!c->err) { c->err = ErrStd(ex.what()); c->result = -1; }
    }
    lk.lock();
    statRunSec_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - r0).count();
    statBatches_++;
    {
      bool seen[6] = {};
      for (CallSlot *c : batch) { statCalls_[c->kind]++; seen[c->kind] = true; }
      for (int k = 0; k < 6; k++) statKindLaunches_[k] += seen[k];
    }
    for (CallSlot *c : batch) {
      c->done = true;
      if (c != &slot) c->cv.notify_one();
    }
  }
  if (!combinePending_.empty()) {            // hand the leadership to a caller that arrived meanwhile
    combinePending_.front()->lead = true;
    combinePending_.front()->cv.notify_one();
  } else {
    combineLeader_ = false;
  }
}

void Engine::RunCombined(const std::vector<CallSlot *> &batch) {
  std::vector<int64_t> ids, args;
  std::vector<CallSlot *> who;
  // ---- StartQuiz: one launch for all new quizzes
  for (CallSlot *c : batch) if (c->kind == 3) who.push_back(c);
  if (!who.empty()) {
    ids.assign(who.size(), -1);
    PqaError *e = StartQuizBatch((int64_t)ids.size(), ids.data());
    for (size_t x = 0; x < who.size(); x++) {
      who[x]->result = e ? -1 : ids[x];
      if (e) who[x]->err = new PqaError(*e);
    }
    delete e;
    ids.clear(); who.clear();
  }
  // ---- NextQuestion: per-call errors come back through ppErrors
  for (CallSlot *c : batch) if (c->kind == 0) { ids.push_back(c->quiz); who.push_back(c); }
  if (!ids.empty()) {
    std::vector<int64_t> out(ids.size(), -1);
    std::vector<void *> errs(ids.size(), nullptr);
    PqaError *e = NextQuestionBatch((int64_t)ids.size(), ids.data(), nullptr, out.data(), errs.data());
    for (size_t x = 0; x < who.size(); x++) {
      who[x]->result = out[x];
      who[x]->err = static_cast<PqaError *>(errs[x]);
      if (e && !who[x]->err) who[x]->err = new PqaError(*e);   // a batch-wide failure (CUDA error) reaches every caller
      if (who[x]->err) who[x]->result = -1;
    }
    delete e;
  }
  // ---- RecordAnswer: validate each call on its own so that one bad call does not fail the others
  ids.clear(); who.clear();
  for (CallSlot *c : batch) if (c->kind == 1) {
    PqaError *e;
    { std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_); e = ValidateRecordAnswer(1, &c->quiz, &c->arg); }
    if (!e && std::find(ids.begin(), ids.end(), c->quiz) != ids.end())   // two clients answering one quiz at once: the second
      e = ErrNoQuizActiveQuestion(c->arg, PQA_FILE_LINE "An attempt to record an answer in a quiz that doesn't"   // one loses, as serially
                                          " have an active question");
    if (e) { c->err = e; continue; }
    ids.push_back(c->quiz); args.push_back(c->arg); who.push_back(c);
  }
  if (!ids.empty()) {
    PqaError *e = RecordAnswerBatch((int64_t)ids.size(), ids.data(), args.data());
    if (e) { for (CallSlot *c : who) c->err = new PqaError(*e); delete e; }
  }
  // ---- ListTopTargets: one launch per distinct maxCount
  std::vector<CallSlot *> tops;
  for (CallSlot *c : batch) if (c->kind == 2) {
    PqaError *e = nullptr;
    if (c->arg < 0) e = ErrNegativeCount(c->arg, PQA_FILE_LINE "|maxCount| must be non-negative.");
    else { std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_); e = CheckQuiz(c->quiz); }
    if (e) { c->err = e; c->result = -1; continue; }
    tops.push_back(c);
  }
  while (!tops.empty()) {
    const int64_t maxCount = tops.front()->arg;
    ids.clear(); who.clear();
    std::vector<CallSlot *> rest;
    for (CallSlot *c : tops) { if (c->arg == maxCount) { ids.push_back(c->quiz); who.push_back(c); } else rest.push_back(c); }
    std::vector<CiRatedTarget> dest((size_t)(ids.size() * std::max<int64_t>(maxCount, 1)));
    std::vector<int64_t> counts(ids.size(), 0);
    PqaError *e = ListTopTargetsBatch((int64_t)ids.size(), ids.data(), maxCount, dest.data(), counts.data());
    for (size_t x = 0; x < who.size(); x++) {
      if (e) { who[x]->err = new PqaError(*e); who[x]->result = -1; continue; }
      who[x]->result = counts[x];
      if (counts[x] > 0) std::memcpy(who[x]->dest, dest.data() + x * maxCount, sizeof(CiRatedTarget) * (size_t)counts[x]);
    }
    delete e;
    tops.swap(rest);
  }
  // ---- RecordQuizTarget: each call validated on its own, then applied in arrival order by one launch
  ids.clear(); args.clear(); who.clear();
  std::vector<double> amounts;
  for (CallSlot *c : batch) if (c->kind == 4) {
    PqaError *e = nullptr;
    {
      std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
      if (c->arg < 0 || c->arg >= T_) e = ErrIndexOutOfRange(c->arg, 0, T_ - 1, PQA_FILE_LINE "Target index is not in KB range.");
      else if (tGaps_.IsGap(c->arg)) e = ErrAbsentId(c->arg, PQA_FILE_LINE "Target index is not in KB (but rather at a gap).");
      else e = CheckQuiz(c->quiz);
    }
    if (e) { c->err = e; continue; }
    ids.push_back(c->quiz); args.push_back(c->arg); amounts.push_back(c->amount); who.push_back(c);
  }
  if (!ids.empty()) {
    PqaError *e = RecordQuizTargetBatch((int64_t)ids.size(), ids.data(), args.data(), amounts.data());
    if (e) { for (CallSlot *c : who) c->err = new PqaError(*e); delete e; }
  }
  // ---- ReleaseQuiz
  for (CallSlot *c : batch) if (c->kind == 5) c->err = ReleaseQuizBatch(1, &c->quiz);
}

int64_t Engine::NextQuestion(PqaError **err, int64_t iQuiz) {
  if (IsSharded()) { int64_t q = -1; *err = NextQuestionBatch(1, &iQuiz, nullptr, &q, nullptr); return -1; }
  if (maintenance_) { *err = WrongMode("compute next question"); return -1; }   // BaseEngine.cpp:422-427
  CallSlot s; s.kind = 0; s.quiz = iQuiz;
  Submit(s);
  *err = s.err;
  return s.err ? -1 : s.result;
}

// ---------------------------------------------------------------------------------------------------------
// RecordAnswer: BaseEngine::RecordAnswer (BaseEngine.cpp:439-464) -> CEQuiz::RecordAnswer (CEQuiz.h:77-122)
PqaError *Engine::RecordAnswerBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pAnswers) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pAnswers");
  if (IsSharded()) return ErrNotImplemented("RecordAnswer on a sharded engine: use PqaB200_ShardRecordAnswerBegin / End");
  if (maintenance_) return WrongMode("record an answer");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return WrongMode("record an answer");   // re-checked under the lock: StartMaintenance may have completed meanwhile
  PQA_TRY
  if (PqaError *e = ValidateRecordAnswer(n, pQuizIds, pAnswers)) return e;
  UploadIds(n, pQuizIds);
  hAnswers_.ensure(n); dAnswers_.ensure(n, stream_);
  std::memcpy(hAnswers_.get(), pAnswers, sizeof(int64_t) * (size_t)n);
  PQA_CU(cudaMemcpyAsync(dAnswers_.get(), hAnswers_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  const int looseW = std::max(1, W_ - 1);  // CpuEngine GetNLooseWorkers, CEQuiz.h:98
  launch_record_answer(kb(), pool(), n, dIds_.get(), dAnswers_.get(), looseW, stream_);
  for (int64_t x = 0; x < n; x++) {
    HostQuiz &q = quizzes_[pQuizIds[x]];
    q.answers.push_back(CiAnsweredQuestion{q.activeQuestion, pAnswers[x]});  // CEQuiz.h:90
    q.activeQuestion = -1;                                                     // :92
  }
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::RecordAnswer(int64_t iQuiz, int64_t iAnswer) {
  if (IsSharded()) return RecordAnswerBatch(1, &iQuiz, &iAnswer);
  if (maintenance_) return WrongMode("record an answer");                       // BaseEngine.cpp:440-444
  CallSlot s; s.kind = 1; s.quiz = iQuiz; s.arg = iAnswer;
  Submit(s);
  return s.err;
}

PqaError *Engine::ValidateRecordAnswer(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) const {
  for (int64_t x = 0; x < n; x++) {  // validate everything before touching any quiz
    if (pAnswers[x] < 0 || pAnswers[x] >= K_)
      return ErrIndexOutOfRange(pAnswers[x], 0, K_ - 1, "Answer index is not in the answer range.");
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
    const HostQuiz &q = quizzes_[pQuizIds[x]];
    if (q.activeQuestion == -1)
      return ErrNoQuizActiveQuestion(pAnswers[x], PQA_FILE_LINE "An attempt to record an answer in a quiz that doesn't"
                                                  " have an active question");
    if (q.activeQuestion < 0 || q.activeQuestion >= Q_)
      return ErrNoQuizActiveQuestion(pAnswers[x], PQA_FILE_LINE "An attempt to record an answer in a quiz that has"
                                                  " invalid active question");
  }
  // The same quiz twice in one launch: the first answer consumes the active question (CEQuiz.h:92), so the second call
  // is what the reference fails with NoQuizActiveQuestion -- and two CTAs must never update one quiz slot concurrently.
  if (n > 1) {
    std::vector<int64_t> sorted(pQuizIds, pQuizIds + n);
    std::sort(sorted.begin(), sorted.end());
    const auto dup = std::adjacent_find(sorted.begin(), sorted.end());
    if (dup != sorted.end())
      return ErrNoQuizActiveQuestion(*dup, PQA_FILE_LINE "An attempt to record an answer in a quiz that doesn't"
                                           " have an active question (the same quiz appears twice in one batch)");
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------------------
// Question-sharded operation: see the protocol in include/PqaB200Ext.h. The buffers are filled here, summed across
// shards by the caller, and consumed by the second half of each operation.
PqaError *Engine::ShardEval(int64_t n, const int64_t *pQuizIds) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (IsTargetSharded()) return ErrNotImplemented("ShardEval on a target-sharded engine: use PqaB200_TShardEvalW / EvalHVL / Priority");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  shardPriorityCount_ = n * Q_;
  dShardPriority_.ensure((size_t)shardPriorityCount_, stream_);
  PQA_CU(cudaMemsetAsync(dShardPriority_.get(), 0, sizeof(double) * (size_t)shardPriorityCount_, stream_));  // +0.0
  EvalDetail det{nullptr, nullptr, nullptr, nullptr};
  launch_eval_questions(kbEval(), pool(), n, dIds_.get(), dShardPriority_.get(), det, evalCfg_, stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));   // the caller's collective runs on another stream
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::ShardSelect(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms, int64_t *pQuestions,
                              void **ppErrors) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (!pQuizIds || !pQuestions || !pRandoms)
    return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pQuestions/pRandoms (every shard must use the same draws)");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (shardPriorityCount_ != n * Q_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "ShardSelect without a matching ShardEval");
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  hRandoms_.ensure(n); dRandoms_.ensure(n, stream_);
  std::memcpy(hRandoms_.get(), pRandoms, sizeof(uint64_t) * (size_t)n);
  PQA_CU(cudaMemcpyAsync(dRandoms_.get(), hRandoms_.get(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  dRunLength_.ensure((size_t)(n * Q_), stream_); dQuestions_.ensure(n, stream_); hQuestions_.ensure(n);
  launch_select_question(kbQuiz(), pool(), n, dIds_.get(), dShardPriority_.get(), dRandoms_.get(), W_, dRunLength_.get(),
                         nullptr, dQuestions_.get(), 1, stream_);
  PQA_CU(cudaMemcpyAsync(hQuestions_.get(), dQuestions_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, stream_));
  QueueAnomalyCopy();
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  ReportAnomalies(hAnom_);
  PqaError *firstErr = nullptr;
  uint64_t nAsked = 0;
  for (int64_t x = 0; x < n; x++) {
    const int64_t qst = hQuestions_.get()[x];
    pQuestions[x] = qst;
    if (ppErrors) ppErrors[x] = nullptr;
    if (qst < 0) {
      PqaError *e = MakeError(ErrCode::QuestionsExhausted, PQA_FILE_LINE "Found no unasked question that is not in a gap.");
      if (ppErrors) ppErrors[x] = e;
      else if (!firstErr) firstErr = e;
      else delete e;
      continue;
    }
    quizzes_[pQuizIds[x]].activeQuestion = qst;
    nAsked++;
  }
  nQuestionsAsked_.fetch_add(nAsked, std::memory_order_relaxed);
  return firstErr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::ShardRecordAnswerBegin(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (!pQuizIds || !pAnswers) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pAnswers");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  if (PqaError *e = ValidateRecordAnswer(n, pQuizIds, pAnswers)) return e;
  UploadIds(n, pQuizIds);
  hAnswers_.ensure(n); dAnswers_.ensure(n, stream_);
  std::memcpy(hAnswers_.get(), pAnswers, sizeof(int64_t) * (size_t)n);
  PQA_CU(cudaMemcpyAsync(dAnswers_.get(), hAnswers_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  const int looseW = std::max(1, W_ - 1);
  shardPriorsCount_ = n * Tp_;
  dShardPriors_.ensure((size_t)shardPriorsCount_, stream_);
  if (IsTargetSharded()) {
    // this shard's columns of m[j] = prior * (sA/mD); the other columns stay +0 for the caller's sum over shards
    PQA_CU(cudaMemsetAsync(dShardPriors_.get(), 0, sizeof(double) * (size_t)shardPriorsCount_, stream_));
    PeerBufs out; out.n = 1; out.p[0] = dShardPriors_.get();
    launch_tshard_record_answer_partial(kb(), pool(), tFirst_, n, dIds_.get(), dAnswers_.get(), out, stream_);
  } else {
    launch_record_answer(kb(), pool(), n, dIds_.get(), dAnswers_.get(), looseW, stream_);   // owner: update; others: zeros
    launch_gather_prior_rows(pool(), n, dIds_.get(), dShardPriors_.get(), stream_);
  }
  for (int64_t x = 0; x < n; x++) {
    HostQuiz &q = quizzes_[pQuizIds[x]];
    q.answers.push_back(CiAnsweredQuestion{q.activeQuestion, pAnswers[x]});
    q.activeQuestion = -1;
  }
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::ShardRecordAnswerEnd(int64_t n, const int64_t *pQuizIds) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (shardPriorsCount_ != n * Tp_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "ShardRecordAnswerEnd without a matching Begin");
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  if (IsTargetSharded())   // complete rows of m[j]: bookkeeping + the reference's normalisation, bit-exact
    launch_tshard_record_answer_finish(kbQuiz(), pool(), n, dIds_.get(), dShardPriors_.get(), std::max(1, W_ - 1), stream_);
  else
    launch_scatter_prior_rows(pool(), n, dIds_.get(), dShardPriors_.get(), stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::ShardBuffer(int32_t which, void **ppDevice, int64_t *pCount) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!ppDevice || !pCount) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "ppDevice/pCount");
  if (which == 0) { *ppDevice = p2pLastPriority_ ? p2pLastPriority_ : dShardPriority_.get(); *pCount = shardPriorityCount_; }
  else if (which == 1) { *ppDevice = dShardPriors_.get(); *pCount = shardPriorsCount_; }
  else if (which == 2) { *ppDevice = dShardW_.get(); *pCount = shardWCount_; }
  else if (which == 3) { *ppDevice = dShardHVL_.get(); *pCount = shardHVLCount_; }
  else return ErrIndexOutOfRange(which, 0, 3, PQA_FILE_LINE "which");
  return nullptr;
}

int64_t Engine::GetActiveQuestionId(PqaError **err, int64_t iQuiz) {
  if (maintenance_) { *err = WrongMode("get active question ID for a quiz"); return -1; }
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  *err = CheckQuiz(iQuiz);
  return *err ? -1 : quizzes_[iQuiz].activeQuestion;
}

PqaError *Engine::SetActiveQuestionBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pQuestions) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pQuestions) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pQuestions");
  if (maintenance_) return WrongMode("get active question ID for a quiz");   // sic, BaseEngine.cpp:491-493
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return WrongMode("get active question ID for a quiz");   // re-checked under the lock: StartMaintenance may have completed meanwhile
  PQA_TRY
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  UploadIds(n, pQuizIds);
  hQuestions_.ensure(n); dQuestions_.ensure(n, stream_);
  std::memcpy(hQuestions_.get(), pQuestions, sizeof(int64_t) * (size_t)n);
  PQA_CU(cudaMemcpyAsync(dQuestions_.get(), hQuestions_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  launch_set_active(pool(), n, dIds_.get(), dQuestions_.get(), stream_);
  for (int64_t x = 0; x < n; x++) quizzes_[pQuizIds[x]].activeQuestion = pQuestions[x];  // BaseQuiz::SetActiveQuestion
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
PqaError *Engine::SetActiveQuestion(int64_t iQuiz, int64_t iQuestion) {
  return SetActiveQuestionBatch(1, &iQuiz, &iQuestion);
}

// ---------------------------------------------------------------------------------------------------------
// ListTopTargets: BaseEngine::ListTopTargets (BaseEngine.cpp:511-527) -> CpuEngine::ListTopTargetsSpec
// (CpuEngine.cpp:417-440). The reference switches to a radix-sort variant when maxCount is a large fraction of
// T/W (:424-432); that variant's merge order is inconsistent in the reference itself (SURVEY.md hard part 9), so
// this engine always runs the heapify algorithm, which is what every top-10 style call reaches.
PqaError *Engine::ListTopTargetsBatch(int64_t n, const int64_t *pQuizIds, int64_t maxCount, CiRatedTarget *pDest,
                                      int64_t *pCounts) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (maxCount < 0) return ErrNegativeCount(maxCount, PQA_FILE_LINE "|maxCount| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pCounts || (maxCount > 0 && !pDest))
    return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pDest/pCounts");
  if (maintenance_) return WrongMode("compute next question");   // sic, BaseEngine.cpp:514-516
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return WrongMode("compute next question");   // re-checked under the lock: StartMaintenance may have completed meanwhile
  PQA_TRY
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  if (maxCount == 0) { for (int64_t x = 0; x < n; x++) pCounts[x] = 0; return nullptr; }
  UploadIds(n, pQuizIds);
  const size_t nItems = (size_t)(n * maxCount);
  dTop_.ensure(nItems, stream_); hTop_.ensure(nItems); dCounts_.ensure(n, stream_); hCounts_.ensure(n);
  const bool needScratch = (size_t)W_ * 32 + (size_t)T_ * 16 > 200 * 1024;
  if (needScratch) dTopScratch_.ensure((size_t)(n * T_), stream_);
  launch_list_top_targets(kbQuiz(), pool(), n, dIds_.get(), W_, maxCount, dTopScratch_.get(), dTop_.get(), dCounts_.get(), stream_);
  PQA_CU(cudaMemcpyAsync(hTop_.get(), dTop_.get(), sizeof(CiRatedTarget) * nItems, cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaMemcpyAsync(hCounts_.get(), dCounts_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  for (int64_t x = 0; x < n; x++) {
    pCounts[x] = hCounts_.get()[x];
    std::memcpy(pDest + x * maxCount, hTop_.get() + x * maxCount, sizeof(CiRatedTarget) * (size_t)pCounts[x]);
  }
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

int64_t Engine::ListTopTargets(PqaError **err, int64_t iQuiz, int64_t maxCount, CiRatedTarget *pDest) {
  if (maxCount > 0 && !pDest) { *err = MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pDest"); return -1; }
  if (maintenance_) { *err = WrongMode("compute next question"); return -1; }   // sic, BaseEngine.cpp:512-517
  CallSlot s; s.kind = 2; s.quiz = iQuiz; s.arg = maxCount; s.dest = pDest;
  Submit(s);
  *err = s.err;
  return s.err ? -1 : s.result;
}

// ---------------------------------------------------------------------------------------------------------
// RecordQuizTarget / Train. The reference applies a quiz' answers two at a time (CpuEngine.cpp:451-462,
// CETrainOperation.cpp:28-83); here every Perform2/Perform1 is split into per-cell operations (same-question pairs
// stay fused because they share mD[q][t]) and grouped by (question, target) so that one device thread applies a
// cell group's operations in the reference's sequence order.
int64_t Engine::AppendQuizOps(TrainOp *dst, const CiAnsweredQuestion *aqs, int64_t n, int64_t iTarget, double amount) {
  int64_t x = 0, w = 0;                    // writes at most n operations, returns how many
  for (; x + 1 < n; x += 2) {
    const CiAnsweredQuestion &f = aqs[x], &s = aqs[x + 1];
    if (f._iQuestion == s._iQuestion) {
      dst[w++] = TrainOp{f._iQuestion, f._iAnswer, s._iAnswer, iTarget, amount};  // doubled step or 3-add form
    } else {
      dst[w++] = TrainOp{f._iQuestion, f._iAnswer, -1, iTarget, amount};
      dst[w++] = TrainOp{s._iQuestion, s._iAnswer, -1, iTarget, amount};
    }
  }
  if (x < n) dst[w++] = TrainOp{aqs[x]._iQuestion, aqs[x]._iAnswer, -1, iTarget, amount};
  return w;
}

// opsAll may live in pinned memory (hOps_, big RecordQuizTarget batches): the upload is then a plain DMA.
PqaError *Engine::ApplyTrain(const TrainOp *opsAll, int64_t nAll, const std::vector<int64_t> &targets,
                             const std::vector<double> &amounts) {
  // caller holds mu_
  PQA_TRY
  std::vector<TrainOp> owned;
  if (qLocal_ != Q_) {   // question-sharded engine: cells of other shards' questions are theirs to update
    for (int64_t x = 0; x < nAll; x++) if (OwnsQuestion(opsAll[x].q)) owned.push_back(opsAll[x]);
  } else if (IsTargetSharded()) {   // target-sharded engine: only the cells of its own columns, addressed locally
    for (int64_t x = 0; x < nAll; x++)
      if (opsAll[x].target >= tFirst_ && opsAll[x].target < tFirst_ + tLocal_) { owned.push_back(opsAll[x]); owned.back().target -= tFirst_; }
  }
  const TrainOp *ops = IsSharded() ? owned.data() : opsAll;
  const int64_t nOps = IsSharded() ? (int64_t)owned.size() : nAll;
  MarkQuestionsChanged(ops, nOps);
  const int64_t nT = (int64_t)targets.size();
  const int64_t kDeviceGrouping = 8192;   // from here on the grouping by cell is a device radix sort (pqa_train_sort.cu)
  if (nOps >= kDeviceGrouping) {
    const size_t need = std::max(train_sort_scratch_bytes(nOps), train_sort_scratch_bytes(nT));
    dSortScratch_.ensure(need, stream_);
    dOps_.ensure(nOps, stream_);
    PQA_CU(cudaMemcpyAsync(dOps_.get(), ops, sizeof(TrainOp) * (size_t)nOps, cudaMemcpyHostToDevice, stream_));
    launch_train_ops_device_grouped(kb(), dOps_.get(), nOps, dSortScratch_.get(), dSortScratch_.size(), stream_);
    if (nT > 0) {
      dTargets_.ensure(nT, stream_); dAmounts_.ensure(nT, stream_);
      PQA_CU(cudaMemcpyAsync(dTargets_.get(), targets.data(), sizeof(int64_t) * (size_t)nT, cudaMemcpyHostToDevice, stream_));
      PQA_CU(cudaMemcpyAsync(dAmounts_.get(), amounts.data(), sizeof(double) * (size_t)nT, cudaMemcpyHostToDevice, stream_));
      launch_add_vb_device_grouped(kbQuiz(), dTargets_.get(), dAmounts_.get(), nT, dSortScratch_.get(), dSortScratch_.size(), stream_);
    }
    PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));   // the host vectors die at scope end
    return nullptr;
  }
  if (nOps > 0) {
    std::vector<int64_t> order(nOps);
    std::iota(order.begin(), order.end(), 0);
    auto key = [&](int64_t o) { return ops[o].q * T_ + ops[o].target; };
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key(a) < key(b); });
    std::vector<TrainOp> sorted(nOps);
    std::vector<int64_t> groupStart;
    for (int64_t x = 0; x < nOps; x++) {
      sorted[x] = ops[order[x]];
      if (x == 0 || key(order[x]) != key(order[x - 1])) groupStart.push_back(x);
    }
    const int64_t nGroups = (int64_t)groupStart.size();
    groupStart.push_back(nOps);
    dOps_.ensure(nOps, stream_); dGroupStart_.ensure(groupStart.size(), stream_);
    PQA_CU(cudaMemcpyAsync(dOps_.get(), sorted.data(), sizeof(TrainOp) * (size_t)nOps, cudaMemcpyHostToDevice, stream_));
    PQA_CU(cudaMemcpyAsync(dGroupStart_.get(), groupStart.data(), sizeof(int64_t) * groupStart.size(), cudaMemcpyHostToDevice, stream_));
    launch_train_ops(kb(), dOps_.get(), dGroupStart_.get(), nGroups, stream_);
    PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));  // the staging vectors die at scope end
  }
  if (nT > 0) {
    std::vector<int64_t> order(nT);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return targets[a] < targets[b]; });
    std::vector<int64_t> st(nT), groupStart;
    std::vector<double> sa(nT);
    for (int64_t x = 0; x < nT; x++) {
      st[x] = targets[order[x]]; sa[x] = amounts[order[x]];
      if (x == 0 || st[x] != st[x - 1]) groupStart.push_back(x);
    }
    const int64_t nGroups = (int64_t)groupStart.size();
    groupStart.push_back(nT);
    dTargets_.ensure(nT, stream_); dAmounts_.ensure(nT, stream_); dGroupStart_.ensure(groupStart.size(), stream_);
    PQA_CU(cudaMemcpyAsync(dTargets_.get(), st.data(), sizeof(int64_t) * (size_t)nT, cudaMemcpyHostToDevice, stream_));
    PQA_CU(cudaMemcpyAsync(dAmounts_.get(), sa.data(), sizeof(double) * (size_t)nT, cudaMemcpyHostToDevice, stream_));
    PQA_CU(cudaMemcpyAsync(dGroupStart_.get(), groupStart.data(), sizeof(int64_t) * groupStart.size(), cudaMemcpyHostToDevice, stream_));
    launch_add_vb(kbQuiz(), dTargets_.get(), dAmounts_.get(), dGroupStart_.get(), nGroups, stream_);
    PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  }
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::RecordQuizTargetBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pTargets,
                                        const double *pAmounts) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pTargets) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pTargets");
  if (pAmounts)
    for (int64_t x = 0; x < n; x++)
      if (!(pAmounts[x] > 0)) return ErrNonPositiveAmount(pAmounts[x], PQA_FILE_LINE "|amount| must be positive.");
  if (maintenance_) return WrongMode("record quiz target");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return WrongMode("record quiz target");   // re-checked under the lock: StartMaintenance may have completed meanwhile
  std::vector<int64_t> targets(n);
  std::vector<double> amounts(n);
  for (int64_t x = 0; x < n; x++) {  // BaseEngine::RecordQuizTarget, BaseEngine.cpp:529-566
    const double amount = pAmounts ? pAmounts[x] : 1.0;
    if (!(amount > 0)) return ErrNonPositiveAmount(amount, PQA_FILE_LINE "|amount| must be positive.");
    if (pTargets[x] < 0 || pTargets[x] >= T_)
      return ErrIndexOutOfRange(pTargets[x], 0, T_ - 1, PQA_FILE_LINE "Target index is not in KB range.");
    if (tGaps_.IsGap(pTargets[x]))
      return ErrAbsentId(pTargets[x], PQA_FILE_LINE "Target index is not in KB (but rather at a gap).");
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
    targets[x] = pTargets[x]; amounts[x] = amount;
  }
  PQA_TRY
  int64_t maxOps = 0;
  for (int64_t x = 0; x < n; x++) maxOps += (int64_t)quizzes_[pQuizIds[x]].answers.size();
  hOps_.ensure((size_t)std::max<int64_t>(maxOps, 1));      // pinned: the operation list is built where the DMA reads it
  int64_t nOps = 0;
  for (int64_t x = 0; x < n; x++) {
    const HostQuiz &q = quizzes_[pQuizIds[x]];
    nOps += AppendQuizOps(hOps_.get() + nOps, q.answers.data(), (int64_t)q.answers.size(), targets[x], amounts[x]);
  }
  return ApplyTrain(hOps_.get(), nOps, targets, amounts);
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::RecordQuizTarget(int64_t iQuiz, int64_t iTarget, double amount) {
  if (IsSharded()) return RecordQuizTargetBatch(1, &iQuiz, &iTarget, &amount);
  if (!(amount > 0)) return ErrNonPositiveAmount(amount, PQA_FILE_LINE "|amount| must be positive.");   // BaseEngine.cpp:530-533
  if (maintenance_) return WrongMode("record quiz target");
  CallSlot s; s.kind = 4; s.quiz = iQuiz; s.arg = iTarget; s.amount = amount;
  Submit(s);
  return s.err;
}

// BaseEngine::Train (BaseEngine.cpp:235-250) -> CpuEngine::TrainSpec (CpuEngine.cpp:102-183): answered questions
// are bucketed by iQuestion % W (CETrainSubtaskDistrib.h:19-52) into LIFO chains that are then consumed two at a
// time (CETrainSubtaskAdd.cpp:17-38). Arrival tickets are taken in array order here (the reference's are racy).
PqaError *Engine::Train(int64_t nQuestions, const CiAnsweredQuestion *pAQs, int64_t iTarget, double amount) {
  if (nQuestions < 0) return ErrNegativeCount(nQuestions, "|nQuestions| must be non-negative.");
  if (!(amount > 0)) return ErrNonPositiveAmount(amount, PQA_FILE_LINE "|amount| must be positive.");
  if (nQuestions > 0 && !pAQs) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pAQs");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (iTarget < 0 || iTarget >= T_) return ErrIndexOutOfRange(iTarget, 0, T_ - 1, "Target index is not in KB range.");
  if (tGaps_.IsGap(iTarget)) return ErrAbsentId(iTarget, PQA_FILE_LINE "Target index is not in KB (but rather at a gap).");   // CpuEngine.cpp:148-152
  for (int64_t x = 0; x < nQuestions; x++) {  // CETrainSubtaskDistrib.h:24-43
    if (pAQs[x]._iQuestion < 0 || pAQs[x]._iQuestion >= Q_)
      return ErrIndexOutOfRange(pAQs[x]._iQuestion, 0, Q_ - 1, PQA_FILE_LINE "Question index is not in KB range.");
    if (qGaps_.IsGap(pAQs[x]._iQuestion))
      return ErrAbsentId(pAQs[x]._iQuestion, "Question index is not in KB (but rather at a gap).");
    if (pAQs[x]._iAnswer < 0 || pAQs[x]._iAnswer >= K_)
      return ErrIndexOutOfRange(pAQs[x]._iAnswer, 0, K_ - 1, PQA_FILE_LINE "Answer index is not in KB range.");
  }
  std::vector<int64_t> last(W_, -1), prev(std::max<int64_t>(nQuestions, 1), -1);
  for (int64_t x = 0; x < nQuestions; x++) {
    const int64_t b = pAQs[x]._iQuestion % W_;
    prev[x] = last[b]; last[b] = x;
  }
  std::vector<TrainOp> ops;
  for (int w = 0; w < W_; w++) {
    int64_t iLast = last[w];
    while (iLast != -1) {
      const CiAnsweredQuestion &f = pAQs[iLast];
      iLast = prev[iLast];
      if (iLast == -1) { ops.push_back(TrainOp{f._iQuestion, f._iAnswer, -1, iTarget, amount}); break; }
      const CiAnsweredQuestion &s = pAQs[iLast];
      iLast = prev[iLast];
      if (f._iQuestion == s._iQuestion) {
        ops.push_back(TrainOp{f._iQuestion, f._iAnswer, s._iAnswer, iTarget, amount});
      } else {
        ops.push_back(TrainOp{f._iQuestion, f._iAnswer, -1, iTarget, amount});
        ops.push_back(TrainOp{s._iQuestion, s._iAnswer, -1, iTarget, amount});
      }
    }
  }
  PqaError *e = ApplyTrain(ops.data(), (int64_t)ops.size(), std::vector<int64_t>{iTarget}, std::vector<double>{amount});
  if (!e) nQuestionsAsked_.fetch_add((uint64_t)nQuestions, std::memory_order_relaxed);  // CpuEngine.cpp:179
  return e;
}

// BaseEngine::ReleaseQuiz (BaseEngine.cpp:568-603)
PqaError *Engine::ReleaseQuizBatch(int64_t n, const int64_t *pQuizIds) {
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n > 0 && !pQuizIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds");
  if (maintenance_) return WrongMode("release quiz");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return WrongMode("release quiz");   // re-checked under the lock: StartMaintenance may have completed meanwhile
  for (int64_t x = 0; x < n; x++) {
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
    HostQuiz &q = quizzes_[pQuizIds[x]];
    q.present = false; q.answers.clear(); q.answers.shrink_to_fit(); q.activeQuestion = -1;
    quizGaps_.push_back(pQuizIds[x]);
    pimQuiz_.RemoveComp(pQuizIds[x]);      // BaseEngine::UnassignQuiz, BaseEngine.cpp:796-801
  }
  if (residentN_ > 0) residentN_ = 0;  // a released quiz may be part of the bound batch
  return nullptr;
}
PqaError *Engine::ReleaseQuiz(int64_t iQuiz) {
  if (IsSharded()) return ReleaseQuizBatch(1, &iQuiz);
  if (maintenance_) return WrongMode("release quiz");
  CallSlot s; s.kind = 5; s.quiz = iQuiz;
  Submit(s);
  return s.err;
}

// ---------------------------------------------------------------------------------------------------------
// IPqaEngine::CopyATargets / CopyDTargets / CopyBTargets (Interface/IPqaEngine.h:36-39, CpuEngine.cpp:690-709)
PqaError *Engine::CopyATargets(int64_t iQuestion, int64_t iAnswer, int64_t maxTargets, double *pFreqs) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (iQuestion < 0 || iQuestion >= Q_) return ErrIndexOutOfRange(iQuestion, 0, Q_ - 1, PQA_FILE_LINE "Question index is not in KB range.");
  if (iAnswer < 0 || iAnswer >= K_) return ErrIndexOutOfRange(iAnswer, 0, K_ - 1, PQA_FILE_LINE "Answer index is not in KB range.");
  if (!OwnsQuestion(iQuestion)) return ErrIndexOutOfRange(iQuestion, qFirst_, qFirst_ + qLocal_ - 1, PQA_FILE_LINE "Question is not in this engine's shard.");
  PQA_TRY
  // a target-sharded engine fills its own columns of pFreqs (positions tFirst .. tFirst+tLocal-1) and leaves the rest
  const int64_t cnt = std::min(maxTargets, tFirst_ + tLocal_) - tFirst_;
  if (cnt > 0) PQA_CU(cudaMemcpy(pFreqs + tFirst_, dSA_ + ((iQuestion - qFirst_) * K_ + iAnswer) * TpL_, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
PqaError *Engine::CopyDTargets(int64_t iQuestion, int64_t maxTargets, double *pFreqs) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (iQuestion < 0 || iQuestion >= Q_) return ErrIndexOutOfRange(iQuestion, 0, Q_ - 1, PQA_FILE_LINE "Question index is not in KB range.");
  if (!OwnsQuestion(iQuestion)) return ErrIndexOutOfRange(iQuestion, qFirst_, qFirst_ + qLocal_ - 1, PQA_FILE_LINE "Question is not in this engine's shard.");
  PQA_TRY
  const int64_t cnt = std::min(maxTargets, tFirst_ + tLocal_) - tFirst_;
  if (cnt > 0) PQA_CU(cudaMemcpy(pFreqs + tFirst_, dMD_ + (iQuestion - qFirst_) * TpL_, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
PqaError *Engine::CopyBTargets(int64_t maxTargets, double *pFreqs) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  const int64_t cnt = std::min(maxTargets, T_);
  if (cnt > 0) PQA_CU(cudaMemcpy(pFreqs, dVB_, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// Whole-KB transfer in the reference's file layout (CpuEngine.cpp:664-688): rows of T doubles, no padding.
PqaError *Engine::UploadKB(const double *sA, const double *mD, const double *vB) {
  if (!sA || !mD || !vB) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "sA/mD/vB");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  // the host arrays always describe the whole KB; a question-sharded engine takes its rows out of them, a
  // target-sharded engine its columns
  sA += qFirst_ * K_ * T_ + tFirst_; mD += qFirst_ * T_ + tFirst_;
  MarkKBChanged();
  const int64_t Q_ = qLocal_;   // rows handled below
  if (Tp_ == T_ && !IsTargetSharded()) {
    PQA_CU(cudaMemcpyAsync(dSA_, sA, sizeof(double) * (size_t)(Q_ * K_ * T_), cudaMemcpyHostToDevice, stream_));
    PQA_CU(cudaMemcpyAsync(dMD_, mD, sizeof(double) * (size_t)(Q_ * T_), cudaMemcpyHostToDevice, stream_));
  } else {
    // stage the rows' local columns through a device scratch in slabs of <= 256 MB, then pad rows on the device
    const int64_t TL = tLocal_;
    const int64_t rowsPerSlab = std::max<int64_t>(1, (256ll << 20) / (TL * 8));
    dRowScratch_.ensure((size_t)(std::min(rowsPerSlab, Q_ * K_) * TL), stream_);
    for (int64_t r0 = 0; r0 < Q_ * K_; r0 += rowsPerSlab) {
      const int64_t nr = std::min(rowsPerSlab, Q_ * K_ - r0);
      PQA_CU(cudaMemcpy2DAsync(dRowScratch_.get(), (size_t)TL * 8, sA + r0 * T_, (size_t)T_ * 8, (size_t)TL * 8, (size_t)nr,
                               cudaMemcpyHostToDevice, stream_));
      launch_pad_rows(dSA_ + r0 * TpL_, dRowScratch_.get(), nr, TL, TpL_, 0.0, stream_);
    }
    for (int64_t r0 = 0; r0 < Q_; r0 += rowsPerSlab) {
      const int64_t nr = std::min(rowsPerSlab, Q_ - r0);
      PQA_CU(cudaMemcpy2DAsync(dRowScratch_.get(), (size_t)TL * 8, mD + r0 * T_, (size_t)T_ * 8, (size_t)TL * 8, (size_t)nr,
                               cudaMemcpyHostToDevice, stream_));
      launch_pad_rows(dMD_ + r0 * TpL_, dRowScratch_.get(), nr, TL, TpL_, 1.0, stream_);
    }
  }
  PQA_CU(cudaMemsetAsync(dVB_, 0, sizeof(double) * (size_t)Tp_, stream_));
  PQA_CU(cudaMemcpyAsync(dVB_, vB, sizeof(double) * (size_t)T_, cudaMemcpyHostToDevice, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::DownloadKB(double *sA, double *mD, double *vB) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  // whole-KB host arrays; a question-sharded engine fills its rows only, a target-sharded engine its columns only
  if (sA) sA += qFirst_ * K_ * T_ + tFirst_;
  if (mD) mD += qFirst_ * T_ + tFirst_;
  const int64_t Q_ = qLocal_;
  const int64_t TL = tLocal_;
  const int64_t rowsPerSlab = std::max<int64_t>(1, (256ll << 20) / (TL * 8));
  for (int64_t r0 = 0; sA && r0 < Q_ * K_; r0 += rowsPerSlab) {
    const int64_t nr = std::min(rowsPerSlab, Q_ * K_ - r0);
    PQA_CU(cudaMemcpy2DAsync(sA + r0 * T_, (size_t)T_ * 8, dSA_ + r0 * TpL_, (size_t)TpL_ * 8, (size_t)TL * 8, (size_t)nr,
                             cudaMemcpyDeviceToHost, stream_));
  }
  for (int64_t r0 = 0; mD && r0 < Q_; r0 += rowsPerSlab) {
    const int64_t nr = std::min(rowsPerSlab, Q_ - r0);
    PQA_CU(cudaMemcpy2DAsync(mD + r0 * T_, (size_t)T_ * 8, dMD_ + r0 * TpL_, (size_t)TpL_ * 8, (size_t)TL * 8, (size_t)nr,
                             cudaMemcpyDeviceToHost, stream_));
  }
  if (vB) PQA_CU(cudaMemcpyAsync(vB, dVB_, sizeof(double) * (size_t)T_, cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::CopyQuizPriors(int64_t iQuiz, double *pPriors) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (PqaError *e = CheckQuiz(iQuiz)) return e;
  PQA_TRY
  PQA_CU(cudaMemcpyAsync(pPriors, dPriors_ + iQuiz * Tp_, sizeof(double) * (size_t)T_, cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
PqaError *Engine::SetQuizPriors(int64_t iQuiz, const double *pPriors) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (PqaError *e = CheckQuiz(iQuiz)) return e;
  PQA_TRY
  PQA_CU(cudaMemcpyAsync(dPriors_ + iQuiz * Tp_, pPriors, sizeof(double) * (size_t)T_, cudaMemcpyHostToDevice, stream_));
  UploadIds(1, &iQuiz);
  launch_refresh_log_priors(pool(), 1, dIds_.get(), stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::SetEvalKernel(int32_t which, int64_t chunkTargets, int64_t quizzesPerCta, int32_t kahanLanesPerThread) {
  if (which < 0 || which > 2) return ErrIndexOutOfRange(which, 0, 2, PQA_FILE_LINE "which");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (kahanLanesPerThread != 0 && kahanLanesPerThread != 1 && kahanLanesPerThread != 2 && kahanLanesPerThread != 4)
    return ErrIndexOutOfRange(kahanLanesPerThread, 0, 4, PQA_FILE_LINE "kahanLanesPerThread must be 0, 1 or 4");
  evalCfg_.which = which; evalCfg_.chunkTargets = chunkTargets; evalCfg_.quizzesPerCta = quizzesPerCta;
  evalCfg_.kahanLanesPerThread = kahanLanesPerThread;
  return nullptr;
}

void Engine::QueueAnomalyCopy() {
  if (dAnom_) cudaMemcpyAsync(hAnom_, dAnom_, sizeof(uint64_t) * kAnomalyKinds, cudaMemcpyDeviceToHost, stream_);
}
// caller holds mu_; `now` = counters read after the stream was synchronised (or published by the fused kernel)
void Engine::ReportAnomalies(const volatile uint64_t *now) {
  static const char *const what[kAnomalyKinds] = {
      "question priorities <= 0 or not finite (CEEvalQsSubtaskConsider.cpp:209-211)",
      "non-finite running totals of the priorities: overflow or underflow in the question evaluation (CpuEngine.cpp:368-371)",
      "grand totals of the priorities <= 0 (CpuEngine.cpp:375-377)"};
  for (int a = 0; a < kAnomalyKinds; a++) {
    const uint64_t v = now[a];
    if (v <= anomSeen_[a]) continue;
    const uint64_t delta = v - anomSeen_[a];
    anomSeen_[a] = v;
    // warn like the reference's logger, but not for every call of a loop: the first 8, then at powers of two
    anomWarnings_++;
    if (anomWarnings_ <= 8 || (anomWarnings_ & (anomWarnings_ - 1)) == 0)
      std::fprintf(stderr, "[probqa_b200] warning: NextQuestion saw %llu %s; %llu so far\n", (unsigned long long)delta,
                   what[a], (unsigned long long)v);
  }
}
PqaError *Engine::AnomalyCounts(uint64_t *pCounts3) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  if (!dAnom_) { for (int a = 0; a < kAnomalyKinds; a++) pCounts3[a] = 0; return nullptr; }
  QueueAnomalyCopy();
  PQA_CU(cudaStreamSynchronize(stream_));
  for (int a = 0; a < kAnomalyKinds; a++) pCounts3[a] = hAnom_[a];
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// Measured at 1000x5x1000 (bench.py e2e, ms per NextQuestion batch): fused kernel 0.05 / 0.16 / 0.18 / 0.24 / 0.38 at 1 / 4 /
// 6 / 8 / 16 quizzes (it streams the derived KB from L2 once per four quizzes); medium-batch kernel + selection + copy
// 0.15 / 0.15 / 0.18 / 0.27 at 6 / 8 / 16 / 32: the fused launch serves batches up to 4 (one tile of it).
bool Engine::UseFewPath(int64_t n) const {
  static const int64_t limit = std::min<int64_t>(eval_few_max(), env_int("PQA_B200_FEW_MAX", 4));
  return n <= limit && !IsSharded() && K_ <= 8 && evalCfg_.which != 1 && evalCfg_.kahanLanesPerThread == 0 &&
         evalCfg_.chunkTargets == 0;
}
void Engine::EnsureFewResources() {
  if (hFew_) return;
  // {questions[kFewHostSlots], sequence word, anomaly counters}
  PQA_CU(cudaHostAlloc((void **)&hFew_, sizeof(int64_t) * (kFewHostSlots + 1 + kAnomalyKinds), cudaHostAllocMapped));
  std::memset((void *)hFew_, 0, sizeof(int64_t) * (kFewHostSlots + 1 + kAnomalyKinds));
  PQA_CU(cudaHostGetDevicePointer(&dFewHost_, (void *)hFew_, 0));
  dFewTickets_.ensure((size_t)eval_few_ticket_count(), stream_);
  PQA_CU(cudaMemsetAsync(dFewTickets_.get(), 0, sizeof(unsigned) * (size_t)eval_few_ticket_count(), stream_));
}

PqaError *Engine::EvalQuestions(int64_t n, const int64_t *pQuizIds, double *pPriorities, double *pRunLength,
                                double *pGrandTotals, int64_t *pnChunks) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (IsTargetSharded()) return ErrNotImplemented("EvalQuestions on a target-sharded engine: use the PqaB200_TShard* protocol");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  const int64_t nChunks = select_chunk_count(Q_, W_);
  if (pnChunks) *pnChunks = nChunks;
  dPriority_.ensure((size_t)(n * Q_), stream_); dRunLength_.ensure((size_t)(n * Q_), stream_);
  dGrand_.ensure((size_t)(n * nChunks), stream_);
  if (UseFewPath(n)) {     // the kernel one-quiz NextQuestion calls run on, evaluation only
    EnsureFewResources();
    const uint64_t noDraws[8] = {};
    dRandoms_.ensure((size_t)n, stream_);     // never read: no question is chosen
    launch_eval_few_select(kbEval(), pool(), (int)n, pQuizIds, noDraws, W_, dPriority_.get(), dRunLength_.get(),
                           dFewTickets_.get(), nullptr, nullptr, 0, dGrand_.get(), dIds_.get(), dRandoms_.get(), stream_);
  } else {
    EvalDetail det{nullptr, nullptr, nullptr, nullptr};
    launch_eval_questions(kbEval(), pool(), n, dIds_.get(), dPriority_.get(), det, evalCfg_, stream_);
    launch_select_question(kbQuiz(), pool(), n, dIds_.get(), dPriority_.get(), nullptr, W_, dRunLength_.get(), dGrand_.get(),
                           nullptr, 0, stream_);
  }
  if (pPriorities) PQA_CU(cudaMemcpyAsync(pPriorities, dPriority_.get(), sizeof(double) * (size_t)(n * Q_), cudaMemcpyDeviceToHost, stream_));
  if (pRunLength) PQA_CU(cudaMemcpyAsync(pRunLength, dRunLength_.get(), sizeof(double) * (size_t)(n * Q_), cudaMemcpyDeviceToHost, stream_));
  if (pGrandTotals) PQA_CU(cudaMemcpyAsync(pGrandTotals, dGrand_.get(), sizeof(double) * (size_t)(n * nChunks), cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::EvalQuestionsDetailed(int64_t iQuiz, double *pW, double *pH, double *pV, double *pLack,
                                        double *pPriorities) {
  return EvalQuestionsDetailedBatch(1, &iQuiz, pW, pH, pV, pLack, pPriorities);
}
// Per-answer metrics of a whole batch, evaluated by whichever kernel the batch size dispatches to (the parity tests use
// this to hold the benched instantiation of the throughput kernel to the oracle: W_k bit for bit).
PqaError *Engine::EvalQuestionsDetailedBatch(int64_t n, const int64_t *pQuizIds, double *pW, double *pH, double *pV,
                                             double *pLack, double *pPriorities) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (IsTargetSharded()) return ErrNotImplemented("EvalQuestionsDetailed on a target-sharded engine");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  const size_t nq = (size_t)(n * Q_), qk = nq * (size_t)K_;
  dDetail_.ensure(3 * qk + nq, stream_); dPriority_.ensure(nq, stream_);
  PQA_CU(cudaMemsetAsync(dDetail_.get(), 0xFF, sizeof(double) * (3 * qk + nq), stream_));  // NaN where not evaluated
  EvalDetail det{dDetail_.get(), dDetail_.get() + qk, dDetail_.get() + 2 * qk, dDetail_.get() + 3 * qk};
  launch_eval_questions(kbEval(), pool(), n, dIds_.get(), dPriority_.get(), det, evalCfg_, stream_);
  if (pW) PQA_CU(cudaMemcpyAsync(pW, det.W, sizeof(double) * qk, cudaMemcpyDeviceToHost, stream_));
  if (pH) PQA_CU(cudaMemcpyAsync(pH, det.H, sizeof(double) * qk, cudaMemcpyDeviceToHost, stream_));
  if (pV) PQA_CU(cudaMemcpyAsync(pV, det.V, sizeof(double) * qk, cudaMemcpyDeviceToHost, stream_));
  if (pLack) PQA_CU(cudaMemcpyAsync(pLack, det.lack, sizeof(double) * nq, cudaMemcpyDeviceToHost, stream_));
  if (pPriorities) PQA_CU(cudaMemcpyAsync(pPriorities, dPriority_.get(), sizeof(double) * nq, cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// ---------------------------------------------------------------------------------------------------------
// Device-resident stepping: the batch's ids and random draws are bound once; a step is one NextQuestion pass
// (evaluation + selection) whose results stay on the device. Used to time the hot path without host traffic.
PqaError *Engine::ResidentBind(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (IsTargetSharded()) return ErrNotImplemented("resident stepping on a target-sharded engine");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  dResIds_.ensure(n, stream_); dResRandoms_.ensure(n, stream_); dResQuestions_.ensure(n, stream_);
  dResPriority_.ensure((size_t)(n * Q_), stream_); dResRunLength_.ensure((size_t)(n * Q_), stream_);
  std::vector<uint64_t> rnd(n);
  for (int64_t x = 0; x < n; x++) rnd[x] = pRandoms ? pRandoms[x] : NextRandom();
  PQA_CU(cudaMemcpyAsync(dResIds_.get(), pQuizIds, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  PQA_CU(cudaMemcpyAsync(dResRandoms_.get(), rnd.data(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  residentN_ = n;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
PqaError *Engine::ResidentStep() {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (residentN_ <= 0) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "no resident batch is bound");
  PQA_TRY
  EvalDetail det{nullptr, nullptr, nullptr, nullptr};
  if (!evEvalStart_) { PQA_CU(cudaEventCreate(&evEvalStart_)); PQA_CU(cudaEventCreate(&evEvalStop_)); }
  PQA_CU(cudaEventRecord(evEvalStart_, stream_));
  launch_eval_questions(kbEval(), pool(), residentN_, dResIds_.get(), dResPriority_.get(), det, evalCfg_, stream_);
  PQA_CU(cudaEventRecord(evEvalStop_, stream_));
  launch_select_question(kbQuiz(), pool(), residentN_, dResIds_.get(), dResPriority_.get(), dResRandoms_.get(), W_,
                         dResRunLength_.get(), nullptr, dResQuestions_.get(), 0, stream_);
  PQA_CU(cudaGetLastError());
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
PqaError *Engine::ResidentFetch(int64_t *pQuestions) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (residentN_ <= 0) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "no resident batch is bound");
  PQA_TRY
  PQA_CU(cudaMemcpyAsync(pQuestions, dResQuestions_.get(), sizeof(int64_t) * (size_t)residentN_, cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
double Engine::ResidentLastEvalMs() {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!evEvalStart_) return -1.0;
  float ms = -1.f;
  if (cudaEventSynchronize(evEvalStop_) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, evEvalStart_, evEvalStop_) != cudaSuccess) return -1.0;
  return (double)ms;
}
PqaError *Engine::Synchronize() {
  DeviceScope devScope(device_);
  PQA_TRY
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  PQA_CU(cudaGetLastError());
  return nullptr;
  PQA_CATCH_RETURN_ERR
}
PqaError *Engine::FlushL2() {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  if (!flushBuf_) {
    flushBytes_ = 256ull << 20;  // > 126 MB L2
    PQA_CU(cudaMalloc(&flushBuf_, flushBytes_));
  }
  launch_flush_l2(flushBuf_, flushBytes_, stream_);
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// ---------------------------------------------------------------------------------------------------------
// Target-sharded evaluation, host-exchanged form (include/PqaB200Ext.h): each call enqueues one phase and waits for it;
// the caller sums buffer 2 (W) / buffer 3 (H, V, lack) over the shards between the calls.
PqaError *Engine::TShardEvalW(int64_t n, const int64_t *pQuizIds) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (!IsTargetSharded()) return ErrNotImplemented("TShardEvalW on an engine without a target shard");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  shardWCount_ = n * Q_ * K_; shardHVLCount_ = 0; shardPriorityCount_ = 0;
  dShardW_.ensure((size_t)shardWCount_, stream_);
  PQA_CU(cudaMemsetAsync(dShardW_.get(), 0, sizeof(double) * (size_t)shardWCount_, stream_));   // asked questions: +0
  PeerBufs out; out.n = 1; out.p[0] = dShardW_.get();
  launch_eval_tshard_w(kbEval(), pool(), tFirst_, n, dIds_.get(), out, evalCfg_, stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::TShardEvalHVL(int64_t n, const int64_t *pQuizIds) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (shardWCount_ != n * Q_ * K_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "TShardEvalHVL without a matching TShardEvalW");
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  shardHVLCount_ = n * Q_ * (2 * K_ + 1);
  dShardHVL_.ensure((size_t)shardHVLCount_, stream_);
  PQA_CU(cudaMemsetAsync(dShardHVL_.get(), 0, sizeof(double) * (size_t)shardHVLCount_, stream_));
  PeerBufs in, out; in.n = 1; in.p[0] = dShardW_.get(); out.n = 1; out.p[0] = dShardHVL_.get();
  launch_eval_tshard_hvl(kbEval(), pool(), tFirst_, n, dIds_.get(), in, out, evalCfg_, stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::TShardPriority(int64_t n, const int64_t *pQuizIds) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (shardWCount_ != n * Q_ * K_ || shardHVLCount_ != n * Q_ * (2 * K_ + 1))
    return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "TShardPriority without matching TShardEvalW / TShardEvalHVL");
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  shardPriorityCount_ = n * Q_;
  dShardPriority_.ensure((size_t)shardPriorityCount_, stream_);
  PeerBufs inW, inHVL; inW.n = 1; inW.p[0] = dShardW_.get(); inHVL.n = 1; inHVL.p[0] = dShardHVL_.get();
  EvalDetail det{nullptr, nullptr, nullptr, nullptr};
  launch_tshard_priority(kb(), pool(), n, dIds_.get(), inW, inHVL, dShardPriority_.get(), det, stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// ---------------------------------------------------------------------------------------------------------
// Shard exchange over peer memory. Inbox layout (identical on every shard; [2] = parity of the operation counter, so a
// shard that runs ahead into the next operation never overwrites what a slower shard is still reading):
//   0    flags[kMaxPeers] u64 (last epoch published by each rank)    64   error flag    128  UUID of the owning GPU (16 bytes)
//   1024 hand-over flags [kP2PMaxTiles] u64 (exact-order pipeline: the previous shard finished tile t of operation e)
//   16384 W-ready flags  [kP2PMaxTiles] u64 (the last shard published the complete W_k of tile t)
//   state [2][cap*Q*K*8]             Kahan lanes handed over by the previous shard               (target shards)
//   W    [2][nRanks][cap*Q*K]        partial normalisers, slot r written by shard r        (target shards)
//   HVL  [2][nRanks][cap*Q*(2K+1)]   partial H/V/lack sums, slot r written by shard r      (target shards)
//   rows [2][cap*Tp]                 RecordAnswer rows: disjoint column slices (target shards) / owner's row (question shards)
//   pri  [2][cap*Q]                  priorities: disjoint question columns                 (question shards)
static const int64_t kP2PMaxTiles = 1920;                         // 15 KB of flag words per array
static const size_t kP2PHeaderBytes = 32768, kP2POffHandOver = 1024, kP2POffWReady = 16384;

PqaError *Engine::P2PSetExactOrder(int32_t on) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!IsTargetSharded()) return ErrNotImplemented("exact-order pipeline on an engine without a target shard (question shards are exact already)");
  if (p2pPending_) return MakeError(ErrCode::WrongMode, PQA_FILE_LINE "a P2P operation is pending");
  p2pExactOrder_ = on != 0;
  return nullptr;
}

PqaError *Engine::P2PInit(int32_t rank, int32_t nRanks, int64_t maxQuizzes, void **ppBase, int64_t *pBytes) {
  if (nRanks < 1 || nRanks > kMaxPeers) return ErrIndexOutOfRange(nRanks, 1, kMaxPeers, PQA_FILE_LINE "nRanks");
  if (rank < 0 || rank >= nRanks) return ErrIndexOutOfRange(rank, 0, nRanks - 1, PQA_FILE_LINE "rank");
  if (maxQuizzes <= 0) return ErrNegativeCount(maxQuizzes, PQA_FILE_LINE "|maxQuizzes| must be positive.");
  if (!IsSharded()) return ErrNotImplemented("P2PInit on an engine without a question or target shard");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (p2pInbox_) return MakeError(ErrCode::WrongMode, PQA_FILE_LINE "the peer-memory inbox exists already");
  PQA_TRY
  PQA_CU(cudaSetDevice(device_));
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const bool ts = IsTargetSharded();
  p2pSzW_ = ts ? up(sizeof(double) * (size_t)(maxQuizzes * Q_ * K_)) : 0;
  p2pSzHVL_ = ts ? up(sizeof(double) * (size_t)(maxQuizzes * Q_ * (2 * K_ + 1))) : 0;
  p2pSzRows_ = up(sizeof(double) * (size_t)(maxQuizzes * Tp_));
  p2pSzPri_ = ts ? 0 : up(sizeof(double) * (size_t)(maxQuizzes * Q_));
  p2pSzState_ = ts ? up(sizeof(double) * (size_t)(maxQuizzes * Q_ * K_ * 8)) : 0;
  p2pOffW_ = kP2PHeaderBytes;
  p2pOffHVL_ = p2pOffW_ + 2 * (size_t)nRanks * p2pSzW_;
  p2pOffRows_ = p2pOffHVL_ + 2 * (size_t)nRanks * p2pSzHVL_;
  p2pOffPri_ = p2pOffRows_ + 2 * p2pSzRows_;
  p2pOffState_ = p2pOffPri_ + 2 * p2pSzPri_;
  p2pBytes_ = p2pOffState_ + 2 * p2pSzState_;
  if (ts) {
    dTileCounters_.ensure((size_t)kP2PMaxTiles, stream_);
    PQA_CU(cudaMemsetAsync(dTileCounters_.get(), 0, sizeof(unsigned) * (size_t)kP2PMaxTiles, stream_));
  }
  PQA_CU(cudaMalloc(&p2pInbox_, p2pBytes_));
  PQA_CU(cudaMemsetAsync(p2pInbox_, 0, p2pBytes_, stream_));
  {   // which physical GPU this inbox lives on: peers compare it with their own (P2PConnect)
    cudaDeviceProp prop;
    PQA_CU(cudaGetDeviceProperties(&prop, device_));
    PQA_CU(cudaMemcpyAsync(p2pInbox_ + 128, &prop.uuid, 16, cudaMemcpyHostToDevice, stream_));
  }
  preload_exchange_kernels((int)K_);
  // size every scratch buffer of the P2P calls now: growing one later would cudaFree, which waits for the whole device
  // -- including another engine's barrier kernel that is waiting for THIS engine (several engines in one process)
  dIds_.ensure((size_t)maxQuizzes, stream_); dRandoms_.ensure((size_t)maxQuizzes, stream_);
  dAnswers_.ensure((size_t)maxQuizzes, stream_); dQuestions_.ensure((size_t)maxQuizzes, stream_);
  dRunLength_.ensure((size_t)(maxQuizzes * Q_), stream_);
  if (ts) dShardPriority_.ensure((size_t)(maxQuizzes * Q_), stream_);
  hIds_.ensure((size_t)maxQuizzes); hRandoms_.ensure((size_t)maxQuizzes); hAnswers_.ensure((size_t)maxQuizzes);
  hQuestions_.ensure((size_t)maxQuizzes);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  p2pRank_ = rank; p2pRanks_ = nRanks; p2pCap_ = maxQuizzes;
  p2pPeer_[rank] = p2pInbox_;
  if (ppBase) *ppBase = p2pInbox_;
  if (pBytes) *pBytes = (int64_t)p2pBytes_;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::P2PExportHandle(uint8_t *pHandle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!pHandle64) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pHandle64");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!p2pInbox_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "P2PInit first");
  PQA_TRY
  cudaIpcMemHandle_t h;
  PQA_CU(cudaIpcGetMemHandle(&h, p2pInbox_));
  std::memcpy(pHandle64, &h, 64);
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::P2POpenHandle(const uint8_t *pHandle64, void **ppPeerBase) {
  if (!pHandle64 || !ppPeerBase) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pHandle64/ppPeerBase");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  PQA_CU(cudaSetDevice(device_));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, pHandle64, 64);
  PQA_CU(cudaIpcOpenMemHandle(ppPeerBase, h, cudaIpcMemLazyEnablePeerAccess));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::P2PConnect(void *const *pBases) {
  if (!pBases) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pBases");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!p2pInbox_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "P2PInit first");
  PQA_TRY
  for (int r = 0; r < p2pRanks_; r++) {
    if (r == p2pRank_) continue;
    if (!pBases[r]) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "a peer inbox base is NULL");
    p2pPeer_[r] = (char *)pBases[r];
    // inboxes of other devices in this process need peer access; IPC-opened ones got it when they were opened
    cudaPointerAttributes at;
    PQA_CU(cudaPointerGetAttributes(&at, pBases[r]));
    // A peer on the same physical GPU (tests; several processes or engines per GPU) changes how the exact-order pipeline
    // waits. The pointer's device ordinal does not tell for an inbox opened through cudaIpc, the GPU's UUID in its header does.
    {
      cudaDeviceProp prop;
      PQA_CU(cudaGetDeviceProperties(&prop, device_));
      char peerUuid[16];
      PQA_CU(cudaMemcpy(peerUuid, (const char *)pBases[r] + 128, 16, cudaMemcpyDeviceToHost));
      if (std::memcmp(peerUuid, &prop.uuid, 16) == 0) p2pSameDevicePeer_ = true;
    }
    if (at.device != device_) {
      PQA_CU(cudaSetDevice(device_));
      const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PQA_CU(e);
      (void)cudaGetLastError();
    }
  }
  p2pConnected_ = true;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

P2PFlags Engine::p2pFlags() const {
  P2PFlags f;
  f.rank = p2pRank_; f.nRanks = p2pRanks_;
  for (int r = 0; r < kMaxPeers; r++) f.flags[r] = (uint64_t *)p2pPeer_[r];
  f.errFlag = (uint64_t *)(p2pInbox_ + 64);
  return f;
}

PqaError *Engine::P2PCheckError() {
  uint64_t flag = 0;
  PQA_TRY
  PQA_CU(cudaMemcpyAsync(&flag, p2pInbox_ + 64, 8, cudaMemcpyDeviceToHost, stream_));
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  if (flag != 0) {
    PQA_CU(cudaMemsetAsync(p2pInbox_ + 64, 0, 8, stream_));        // reported once; the caller decides whether to go on
    return MakeError(ErrCode::Internal, PQA_FILE_LINE "peer-memory barrier timed out: a shard did not reach epoch " +
                     std::to_string(flag) + " (all shards must issue the same P2P calls in the same order)");
  }
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

static const uint64_t kP2PTimeoutNs = 20ull * 1000 * 1000 * 1000;

PqaError *Engine::P2PNextQuestionBegin(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (!pQuizIds || !pRandoms)
    return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pRandoms (every shard must use the same draws)");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!p2pConnected_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "P2PInit / P2PConnect first");
  if (p2pPending_) return MakeError(ErrCode::WrongMode, PQA_FILE_LINE "a P2P operation is pending: call its End first");
  if (n > p2pCap_) return ErrIndexOutOfRange(n, 1, p2pCap_, PQA_FILE_LINE "more quizzes than the inbox was sized for");
  if (!IsTargetSharded() && evalCfg_.which == 1) return ErrNotImplemented("peer-memory exchange with the exact kernel");
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PQA_TRY
  UploadIds(n, pQuizIds);
  hRandoms_.ensure(n); dRandoms_.ensure(n, stream_);
  std::memcpy(hRandoms_.get(), pRandoms, sizeof(uint64_t) * (size_t)n);
  PQA_CU(cudaMemcpyAsync(dRandoms_.get(), hRandoms_.get(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  dRunLength_.ensure((size_t)(n * Q_), stream_); dQuestions_.ensure(n, stream_); hQuestions_.ensure(n);
  const size_t par = (size_t)(p2pOps_++ & 1);
  const P2PFlags flags = p2pFlags();
  double *priority = nullptr;
  if (IsTargetSharded()) {
    PeerBufs outW, inW, outHVL, inHVL;
    outW.n = inW.n = outHVL.n = inHVL.n = p2pRanks_;
    for (int r = 0; r < p2pRanks_; r++) {
      outW.p[r] = (double *)(p2pPeer_[r] + p2pOffW_ + (par * p2pRanks_ + p2pRank_) * p2pSzW_);
      inW.p[r] = (double *)(p2pInbox_ + p2pOffW_ + (par * p2pRanks_ + r) * p2pSzW_);
      outHVL.p[r] = (double *)(p2pPeer_[r] + p2pOffHVL_ + (par * p2pRanks_ + p2pRank_) * p2pSzHVL_);
      inHVL.p[r] = (double *)(p2pInbox_ + p2pOffHVL_ + (par * p2pRanks_ + r) * p2pSzHVL_);
    }
    const DeviceKB kbE = kbEval();       // brings the derived KB up to date before the first timed phase
    if (!p2pEv_[0]) for (cudaEvent_t &e : p2pEv_) PQA_CU(cudaEventCreate(&e));
    PQA_CU(cudaEventRecord(p2pEv_[0], stream_));
    if (!p2pExactOrder_ || p2pRanks_ == 1) {
      launch_eval_tshard_w(kbE, pool(), tFirst_, n, dIds_.get(), outW, evalCfg_, stream_);   // partial W_k -> every inbox
      PQA_CU(cudaEventRecord(p2pEv_[1], stream_));
      launch_p2p_barrier(flags, ++p2pEpoch_, kP2PTimeoutNs, stream_);
      PQA_CU(cudaEventRecord(p2pEv_[2], stream_));
      launch_eval_tshard_hvl(kbE, pool(), tFirst_, n, dIds_.get(), inW, outHVL, evalCfg_, stream_);
    } else {
      // Exact-order pipeline: the Kahan lanes travel shard 0 -> 1 -> ... -> N-1 tile by tile (tiles of consecutive
      // questions, about two waves of CTAs each); shard N-1 finishes the reference's own sum and publishes W_k to every
      // shard, whose phase-2 CTAs wait per tile. All waiting and signalling happens inside the two kernels.
      const uint64_t opEpoch = p2pOps_;                      // identical on all shards, grows with every operation
      const int64_t gridY = tshard_quiz_tiles(n);
      int64_t tileQ = std::max<int64_t>(1, (2 * (int64_t)smCount_ + gridY - 1) / gridY);
      if (const int64_t forced = env_int("PQA_B200_PIPE_TILE", 0)) tileQ = forced;      // experiments: questions per pipeline tile
      if ((Q_ + tileQ - 1) / tileQ > kP2PMaxTiles) tileQ = (Q_ + kP2PMaxTiles - 1) / kP2PMaxTiles;
      const bool first = p2pRank_ == 0, last = p2pRank_ == p2pRanks_ - 1;
      PipeCtl p1;
      p1.tileQ = tileQ; p1.epoch = opEpoch; p1.timeoutNs = kP2PTimeoutNs; p1.errFlag = (uint64_t *)(p2pInbox_ + 64);
      p1.tileCounters = dTileCounters_.get();
      p1.waitFlags = first ? nullptr : (const uint64_t *)(p2pInbox_ + kP2POffHandOver);
      if (last) {
        for (int r = 0; r < p2pRanks_; r++) p1.signalFlags[p1.nSignal++] = (uint64_t *)(p2pPeer_[r] + kP2POffWReady);
      } else {
        p1.signalFlags[p1.nSignal++] = (uint64_t *)(p2pPeer_[p2pRank_ + 1] + kP2POffHandOver);
      }
      const double *inState = first ? nullptr : (const double *)(p2pInbox_ + p2pOffState_ + par * p2pSzState_);
      double *outState = last ? nullptr : (double *)(p2pPeer_[p2pRank_ + 1] + p2pOffState_ + par * p2pSzState_);
      PeerBufs wAll;                                        // the complete W_k lives in slot 0 of every inbox
      if (last) { wAll.n = p2pRanks_; for (int r = 0; r < p2pRanks_; r++) wAll.p[r] = (double *)(p2pPeer_[r] + p2pOffW_ + (par * p2pRanks_) * p2pSzW_); }
      if (p2pSameDevicePeer_ && !first) {
        // several shard engines share this GPU (tests): phase-1 CTAs that spin for the previous shard's hand-over could fill
        // every SM slot before that shard's own CTAs are resident, so a one-CTA kernel waits for all tiles instead
        const int64_t nTiles = (Q_ + tileQ - 1) / tileQ;
        launch_p2p_wait(p1.waitFlags, (int)nTiles, opEpoch, p1.errFlag, kP2PTimeoutNs, stream_);
        p1.waitFlags = nullptr;
      }
      launch_eval_tshard_w(kbE, pool(), tFirst_, n, dIds_.get(), wAll, evalCfg_, stream_, inState, outState, &p1);
      PQA_CU(cudaEventRecord(p2pEv_[1], stream_));
      PipeCtl p2;
      p2.tileQ = tileQ; p2.epoch = opEpoch; p2.timeoutNs = kP2PTimeoutNs; p2.errFlag = p1.errFlag;
      p2.waitFlags = (const uint64_t *)(p2pInbox_ + kP2POffWReady);
      inW.n = 1;                                            // slot 0 = the complete W_k (inHVL still sums all shards)
      if (p2pSameDevicePeer_) {
        // several shard engines share this GPU (tests): phase-2 CTAs that spin for W_k would occupy the SMs the other
        // engines' phase 1 needs, so a one-CTA kernel waits for all tiles instead
        const int64_t nTiles = (Q_ + tileQ - 1) / tileQ;
        launch_p2p_wait(p2.waitFlags, (int)nTiles, opEpoch, p1.errFlag, kP2PTimeoutNs, stream_);
        p2.waitFlags = nullptr;
      }
      PQA_CU(cudaEventRecord(p2pEv_[2], stream_));
      launch_eval_tshard_hvl(kbE, pool(), tFirst_, n, dIds_.get(), inW, outHVL, evalCfg_, stream_, &p2);
    }
    PQA_CU(cudaEventRecord(p2pEv_[3], stream_));
    launch_p2p_barrier(flags, ++p2pEpoch_, kP2PTimeoutNs, stream_);
    PQA_CU(cudaEventRecord(p2pEv_[4], stream_));
    dShardPriority_.ensure((size_t)(n * Q_), stream_);
    priority = dShardPriority_.get();
    EvalDetail det{nullptr, nullptr, nullptr, nullptr};
    launch_tshard_priority(kb(), pool(), n, dIds_.get(), inW, inHVL, priority, det, stream_);
    shardWCount_ = 0; shardHVLCount_ = 0;
    p2pLastPriority_ = nullptr;
  } else {
    priority = (double *)(p2pInbox_ + p2pOffPri_ + par * p2pSzPri_);
    EvalConfig cfg = evalCfg_;
    for (int r = 0; r < p2pRanks_; r++)
      if (r != p2pRank_) cfg.mirror.p[cfg.mirror.n++] = (double *)(p2pPeer_[r] + p2pOffPri_ + par * p2pSzPri_);
    EvalDetail det{nullptr, nullptr, nullptr, nullptr};
    launch_eval_questions(kbEval(), pool(), n, dIds_.get(), priority, det, cfg, stream_);   // own columns -> every inbox
    launch_p2p_barrier(flags, ++p2pEpoch_, kP2PTimeoutNs, stream_);
    p2pLastPriority_ = priority;
  }
  shardPriorityCount_ = n * Q_;
  launch_select_question(kbQuiz(), pool(), n, dIds_.get(), priority, dRandoms_.get(), W_, dRunLength_.get(), nullptr,
                         dQuestions_.get(), 1, stream_);
  if (IsTargetSharded()) PQA_CU(cudaEventRecord(p2pEv_[5], stream_));
  p2pPhasesValid_ = IsTargetSharded();
  PQA_CU(cudaMemcpyAsync(hQuestions_.get(), dQuestions_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, stream_));
  QueueAnomalyCopy();
  PQA_CU(cudaGetLastError());      // a kernel that failed to launch would leave the peers waiting at the barrier
  p2pPending_ = true;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::P2PNextQuestionEnd(int64_t n, const int64_t *pQuizIds, int64_t *pQuestions, void **ppErrors) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (!pQuizIds || !pQuestions) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pQuestions");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!p2pPending_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "P2PNextQuestionEnd without a Begin");
  p2pPending_ = false;
  if (PqaError *e = P2PCheckError()) return e;
  if (p2pRank_ == 0) ReportAnomalies(hAnom_);      // every shard counts the same; one of them speaks
  PqaError *firstErr = nullptr;
  uint64_t nAsked = 0;
  for (int64_t x = 0; x < n; x++) {
    const int64_t qst = hQuestions_.get()[x];
    pQuestions[x] = qst;
    if (ppErrors) ppErrors[x] = nullptr;
    if (qst < 0) {
      PqaError *e = MakeError(ErrCode::QuestionsExhausted, PQA_FILE_LINE "Found no unasked question that is not in a gap.");
      if (ppErrors) ppErrors[x] = e;
      else if (!firstErr) firstErr = e;
      else delete e;
      continue;
    }
    quizzes_[pQuizIds[x]].activeQuestion = qst;
    nAsked++;
  }
  nQuestionsAsked_.fetch_add(nAsked, std::memory_order_relaxed);
  return firstErr;
}

// Device time (ms) of the five stages of the most recent target-sharded P2PNextQuestion: phase 1 (W_k partials), the
// exchange barrier, phase 2 (H/V/lack partials; with the exact-order pipeline it includes waiting for the tiles' W_k), the
// second barrier, epilogue + selection. From CUDA events on this engine's stream; call after the End.
PqaError *Engine::P2PLastPhaseMs(double *pMs5) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!p2pPhasesValid_ || p2pPending_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "no finished target-sharded P2PNextQuestion");
  PQA_TRY
  PQA_CU(cudaEventSynchronize(p2pEv_[5]));
  for (int x = 0; x < 5; x++) {
    float ms = 0.f;
    PQA_CU(cudaEventElapsedTime(&ms, p2pEv_[x], p2pEv_[x + 1]));
    pMs5[x] = (double)ms;
  }
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::P2PRecordAnswerBegin(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) {
  if (n <= 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be positive.");
  if (!pQuizIds || !pAnswers) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pAnswers");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!p2pConnected_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "P2PInit / P2PConnect first");
  if (p2pPending_) return MakeError(ErrCode::WrongMode, PQA_FILE_LINE "a P2P operation is pending: call its End first");
  if (n > p2pCap_) return ErrIndexOutOfRange(n, 1, p2pCap_, PQA_FILE_LINE "more quizzes than the inbox was sized for");
  PQA_TRY
  if (PqaError *e = ValidateRecordAnswer(n, pQuizIds, pAnswers)) return e;
  UploadIds(n, pQuizIds);
  hAnswers_.ensure(n); dAnswers_.ensure(n, stream_);
  std::memcpy(hAnswers_.get(), pAnswers, sizeof(int64_t) * (size_t)n);
  PQA_CU(cudaMemcpyAsync(dAnswers_.get(), hAnswers_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
  const int looseW = std::max(1, W_ - 1);
  const size_t par = (size_t)(p2pOps_++ & 1);
  const P2PFlags flags = p2pFlags();
  double *rows = (double *)(p2pInbox_ + p2pOffRows_ + par * p2pSzRows_);
  PeerBufs out;
  if (IsTargetSharded()) {
    out.n = p2pRanks_;
    for (int r = 0; r < p2pRanks_; r++) out.p[r] = (double *)(p2pPeer_[r] + p2pOffRows_ + par * p2pSzRows_);
    launch_tshard_record_answer_partial(kb(), pool(), tFirst_, n, dIds_.get(), dAnswers_.get(), out, stream_);
    launch_p2p_barrier(flags, ++p2pEpoch_, kP2PTimeoutNs, stream_);
    launch_tshard_record_answer_finish(kbQuiz(), pool(), n, dIds_.get(), rows, looseW, stream_);
  } else {
    hQuestions_.ensure(n); dQuestions_.ensure(n, stream_);
    for (int64_t x = 0; x < n; x++) hQuestions_.get()[x] = quizzes_[pQuizIds[x]].activeQuestion;
    PQA_CU(cudaMemcpyAsync(dQuestions_.get(), hQuestions_.get(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice, stream_));
    launch_record_answer(kb(), pool(), n, dIds_.get(), dAnswers_.get(), looseW, stream_);   // owner: update; others: bookkeeping
    for (int r = 0; r < p2pRanks_; r++)
      if (r != p2pRank_) out.p[out.n++] = (double *)(p2pPeer_[r] + p2pOffRows_ + par * p2pSzRows_);
    launch_p2p_push_prior_rows(kb(), pool(), n, dIds_.get(), dQuestions_.get(), out, stream_);
    launch_p2p_barrier(flags, ++p2pEpoch_, kP2PTimeoutNs, stream_);
    launch_p2p_pull_prior_rows(kb(), pool(), n, dIds_.get(), dQuestions_.get(), rows, stream_);
  }
  for (int64_t x = 0; x < n; x++) {
    HostQuiz &q = quizzes_[pQuizIds[x]];
    q.answers.push_back(CiAnsweredQuestion{q.activeQuestion, pAnswers[x]});
    q.activeQuestion = -1;
  }
  PQA_CU(cudaGetLastError());
  p2pPending_ = true;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *Engine::P2PRecordAnswerEnd() {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!p2pPending_) return MakeError(ErrCode::NotInitialized, PQA_FILE_LINE "P2PRecordAnswerEnd without a Begin");
  p2pPending_ = false;
  return P2PCheckError();
}

PqaError *Engine::FillBinarySearchKB(double rounds) {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  MarkKBChanged();
  launch_fill_binary_search_kb(kb(), tFirst_, T_, initAmount_, rounds, stream_);
  PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// ---------------------------------------------------------------------------------------------------------
// KB file, byte-compatible with the reference (BaseEngine::LockedSaveKB BaseEngine.cpp:323-385, CpuEngine::SaveStatistics
// CpuEngine.cpp:664-688, BaseEngine::WriteGaps :142-152, PermanentIdManager::Save PermanentIdManager.cpp:27-39; load side
// PqaEngineBaseFactory::LoadEngineDefinition PqaEngineBaseFactory.cpp:56-83, BaseEngine ctor BaseEngine.cpp:26-57):
//   u64 PrecisionDefinition {type:4, mantissa:28, exponent:16, reserved:16}; i64 nAnswers, nQuestions, nTargets;
//   u64 nQuestionsAsked; sA as Q*K rows of T doubles; mD as Q rows; vB;
//   question gaps {i64 n, n x i64}; target gaps; id maps of questions, targets {i64 nextPermId, i64 nComp, nComp x i64};
//   quiz id map saved empty {nextPermId, 0}.
namespace {
struct FileCloser { FILE *f; ~FileCloser() { if (f) std::fclose(f); } };
bool wr(FILE *f, const void *p, size_t bytes) { return std::fwrite(p, 1, bytes, f) == bytes; }
bool rd(FILE *f, void *p, size_t bytes) { return std::fread(p, 1, bytes, f) == bytes; }
PqaError *FileOpErr(const char *path, const std::string &what) {
  return MakeError(ErrCode::FileOp, what, std::string("filePath=[") + path + "]");
}
}  // namespace

// The cells of this engine (all of them, or its question rows / target columns) go to their places in the file: sA at 40,
// mD after it, vB after that (CpuEngine::SaveStatistics, CpuEngine.cpp:664-688). `frame` = also write the header, vB and the
// tail (gap lists, id maps); a sharded KB is saved by one shard writing the frame first and then every shard writing its
// cells into the same file (PqaB200_SaveKBShard).
PqaError *Engine::WriteKBFile(FILE *f, const char *filePath, bool frame) {
  const int64_t hdr = 40;
  const int64_t offA = hdr, offD = offA + Q_ * K_ * T_ * 8, offB = offD + Q_ * T_ * 8, offTail = offB + T_ * 8;
  if (frame) {
    const uint64_t prec = (uint64_t)3 | ((uint64_t)(precMantissa_ & 0xFFFFFFF) << 4) | ((uint64_t)(precExponent_ & 0xFFFF) << 32);
    const int64_t dims[3] = {K_, Q_, T_};
    const uint64_t asked = nQuestionsAsked_.load(std::memory_order_acquire);
    if (std::fseek(f, 0, SEEK_SET) != 0 || !wr(f, &prec, 8) || !wr(f, dims, 24) || !wr(f, &asked, 8))
      return FileOpErr(filePath, PQA_FILE_LINE "Can't write the KB header.");
  }
  // stream the cells out in slabs of <= 64 MB of (local) rows
  const int64_t TL = tLocal_;
  const int64_t rowsPerSlab = std::max<int64_t>(1, (64ll << 20) / (TL * 8));
  std::vector<double> slab((size_t)(rowsPerSlab * TL));
  auto dumpRows = [&](const double *dBase, int64_t stride, int64_t nRows, int64_t nCols, int64_t fileOff, int64_t fileRowDoubles) -> PqaError * {
    for (int64_t r0 = 0; r0 < nRows; r0 += rowsPerSlab) {
      const int64_t nr = std::min(rowsPerSlab, nRows - r0);
      try {
        PQA_CU(cudaMemcpy2DAsync(slab.data(), (size_t)nCols * 8, dBase + r0 * stride, (size_t)stride * 8, (size_t)nCols * 8, (size_t)nr,
                                 cudaMemcpyDeviceToHost, stream_));
        PQA_CU(cudaGetLastError()); PQA_CU(cudaStreamSynchronize(stream_));
      } catch (const CudaFail &cf) { return ErrCuda(cf.code, cf.what(), cf.file, cf.line); }
      if (nCols == fileRowDoubles) {            // whole rows: one contiguous write
        if (fseeko(f, (off_t)(fileOff + r0 * fileRowDoubles * 8), SEEK_SET) != 0 || !wr(f, slab.data(), (size_t)(nr * nCols) * 8))
          return FileOpErr(filePath, PQA_FILE_LINE "Can't write KB rows.");
      } else {                                  // a column slice of every row
        for (int64_t r = 0; r < nr; r++)
          if (fseeko(f, (off_t)(fileOff + (r0 + r) * fileRowDoubles * 8), SEEK_SET) != 0 || !wr(f, slab.data() + r * nCols, (size_t)nCols * 8))
            return FileOpErr(filePath, PQA_FILE_LINE "Can't write KB rows.");
      }
    }
    return nullptr;
  };
  if (PqaError *e = dumpRows(dSA_, TpL_, qLocal_ * K_, TL, offA + (qFirst_ * K_ * T_ + tFirst_) * 8, T_)) return e;
  if (PqaError *e = dumpRows(dMD_, TpL_, qLocal_, TL, offD + (qFirst_ * T_ + tFirst_) * 8, T_)) return e;
  if (frame) {
    if (PqaError *e = dumpRows(dVB_, Tp_, 1, T_, offB, T_)) return e;
    if (fseeko(f, (off_t)offTail, SEEK_SET) != 0) return FileOpErr(filePath, PQA_FILE_LINE "Can't seek to the KB tail.");
    auto dumpGaps = [&](const GapSet &g) {   // BaseEngine::WriteGaps, BaseEngine.cpp:142-152: count, then the ids in LIFO order
      const int64_t nGaps = g.GetNGaps();
      return wr(f, &nGaps, 8) && (nGaps == 0 || wr(f, g.Gaps().data(), (size_t)nGaps * 8));
    };
    if (!dumpGaps(qGaps_)) return FileOpErr(filePath, PQA_FILE_LINE "Can't write the question gaps.");
    if (!dumpGaps(tGaps_)) return FileOpErr(filePath, PQA_FILE_LINE "Can't write the target gaps.");
    if (!pimQ_.Save(f)) return FileOpErr(filePath, PQA_FILE_LINE "Can't write the question permanent-compact ID mappings.");
    if (!pimT_.Save(f)) return FileOpErr(filePath, PQA_FILE_LINE "Can't write the target permanent-compact ID mappings.");
    if (!pimQuiz_.Save(f, true)) return FileOpErr(filePath, PQA_FILE_LINE "Can't write the quiz permanent-compact ID mappings.");
  }
  if (std::fflush(f) != 0) return FileOpErr(filePath, PQA_FILE_LINE "Can't flush the KB file.");
  return nullptr;
}

PqaError *Engine::SaveKB(const char *filePath) {
  if (!filePath) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "Nullptr is passed in place of KB file name.");
  if (IsSharded()) return ErrNotImplemented("SaveKB on a sharded engine: use PqaB200_SaveKBShard on every shard");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);   // the KB must not be trained while it is written (reference: shared lock)
  FileCloser fc{std::fopen(filePath, "wb")};
  if (!fc.f) return MakeError(ErrCode::CantOpenFile, PQA_FILE_LINE "Can't open the KB file to write.",
                              std::string("filePath=[") + filePath + "]");
  return WriteKBFile(fc.f, filePath, true);
}

PqaError *Engine::Shutdown(const char *saveFilePath) {
  if (shutdown_.exchange(true, std::memory_order_acq_rel))          // MaintenanceSwitch::Shutdown (MaintenanceSwitch.cpp:82-90)
    return MakeError(ErrCode::ObjectShutDown, std::string("MaintenanceSwitch seems already shut down.") +
                     (saveFilePath ? std::string(" Not saving file: ") + saveFilePath : std::string()), "CpuEngine<taNumber>::Shutdown()");
  PqaError *err = nullptr;
  if (saveFilePath && *saveFilePath) err = IsSharded() ? SaveKBShard(saveFilePath, true) : SaveKB(saveFilePath);
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t i = 0; i < (int64_t)quizzes_.size(); i++)           // BaseEngine.cpp:289-303
    if (quizzes_[(size_t)i].present) pimQuiz_.RemoveComp(i);
  quizzes_.clear(); quizGaps_.clear();
  pimQuiz_.OnCompact(0, nullptr);
  ReleaseDeviceState();                                            // :311 DestroyStatistics
  return err;
}
void Engine::ReleaseDeviceState() {
  if (stream_) cudaStreamSynchronize(stream_);
  cudaFree(dSA_); cudaFree(dMD_); cudaFree(dVB_); cudaFree(dDerR_);
  dSA_ = dMD_ = dVB_ = dDerR_ = dDerL_ = nullptr; derCapR_ = derCapL_ = 0;
  cudaFree(dPriors_); cudaFree(dLogPriors_); cudaFree(dAsked_); cudaFree(dActive_); cudaFree(dNormS_);
  dPriors_ = dLogPriors_ = dNormS_ = nullptr; dAsked_ = nullptr; dActive_ = nullptr;
  quizCap_ = 0; residentN_ = 0;
}

// One shard's part of a KB file shared by all shards. The shard called with writeFrame != 0 goes first (it creates the
// file and writes header, vB and tail); the others then open the existing file and write their cells in place.
PqaError *Engine::SaveKBShard(const char *filePath, bool writeFrame) {
  if (!filePath) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "Nullptr is passed in place of KB file name.");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  FileCloser fc{std::fopen(filePath, writeFrame ? "wb" : "r+b")};
  if (!fc.f) return MakeError(ErrCode::CantOpenFile, PQA_FILE_LINE "Can't open the KB file to write.",
                              std::string("filePath=[") + filePath + "]");
  return WriteKBFile(fc.f, filePath, writeFrame);
}

Engine *Engine::LoadKB(const char *filePath, const CiB200Options &opts, PqaError **err) {
  *err = nullptr;
  if (!filePath) { *err = MakeError(ErrCode::NullArgument, PQA_FILE_LINE "Nullptr is passed in place of KB file name."); return nullptr; }
  FileCloser fc{std::fopen(filePath, "rb")};
  if (!fc.f) {
    *err = MakeError(ErrCode::CantOpenFile, PQA_FILE_LINE "Can't open the KB file to read.", std::string("filePath=[") + filePath + "]");
    return nullptr;
  }
  uint64_t prec = 0, asked = 0;
  int64_t dims[3] = {0, 0, 0};
  if (!rd(fc.f, &prec, 8) || !rd(fc.f, dims, 24) || !rd(fc.f, &asked, 8)) { *err = FileOpErr(filePath, PQA_FILE_LINE "Can't read the KB header."); return nullptr; }
  if ((prec & 0xF) != 3) { *err = ErrNotImplemented("B200 engine on precision type other than double (KB file header)."); return nullptr; }
  if (dims[0] < 2 || dims[1] < 1 || dims[2] < 2) { *err = ErrInsufficientDims(dims[0], dims[1], dims[2]); return nullptr; }
  CiEngineDefinition def;
  std::memset(&def, 0, sizeof(def));
  def._nAnswers = dims[0]; def._nQuestions = dims[1]; def._nTargets = dims[2];
  def._precType = 3; def._precMantissa = (uint32_t)((prec >> 4) & 0xFFFFFFF); def._precExponent = (uint16_t)((prec >> 32) & 0xFFFF);
  def._initAmount = 1.0;   // not stored in the file; only used to fill a fresh KB, which is overwritten below
  const int64_t K = dims[0], Q = dims[1], T = dims[2];
  std::unique_ptr<Engine> eng(new Engine(def, opts));
  // Stream this engine's cells (all rows, or the question rows / target columns of its shard) from the file to the device
  // in slabs of <= 64 MB; the whole KB is never held on the host.
  const int64_t offA = 40, offD = offA + Q * K * T * 8, offB = offD + Q * T * 8, offTail = offB + T * 8;
  {
    Engine &E = *eng;
    const int64_t TL = E.tLocal_;
    const int64_t rowsPerSlab = std::max<int64_t>(1, (64ll << 20) / (TL * 8));
    PinBuf<double> slab;
    slab.ensure((size_t)(rowsPerSlab * TL));
    auto loadRows = [&](double *dBase, int64_t stride, int64_t nRows, int64_t nCols, int64_t fileOff, int64_t fileRowDoubles) -> bool {
      for (int64_t r0 = 0; r0 < nRows; r0 += rowsPerSlab) {
        const int64_t nr = std::min(rowsPerSlab, nRows - r0);
        if (nCols == fileRowDoubles) {
          if (fseeko(fc.f, (off_t)(fileOff + r0 * fileRowDoubles * 8), SEEK_SET) != 0 || !rd(fc.f, slab.get(), (size_t)(nr * nCols) * 8)) return false;
        } else {
          for (int64_t r = 0; r < nr; r++)
            if (fseeko(fc.f, (off_t)(fileOff + (r0 + r) * fileRowDoubles * 8), SEEK_SET) != 0 || !rd(fc.f, slab.get() + r * nCols, (size_t)nCols * 8)) return false;
        }
        if (cudaMemcpy2DAsync(dBase + r0 * stride, (size_t)stride * 8, slab.get(), (size_t)nCols * 8, (size_t)nCols * 8, (size_t)nr,
                              cudaMemcpyHostToDevice, E.stream_) != cudaSuccess || cudaStreamSynchronize(E.stream_) != cudaSuccess) return false;
      }
      return true;
    };
    // rows are padded on the device: the fill of the constructor left the padding lanes at their neutral values
    if (!loadRows(E.dSA_, E.TpL_, E.qLocal_ * K, TL, offA + (E.qFirst_ * K * T + E.tFirst_) * 8, T) ||
        !loadRows(E.dMD_, E.TpL_, E.qLocal_, TL, offD + (E.qFirst_ * T + E.tFirst_) * 8, T)) {
      *err = FileOpErr(filePath, PQA_FILE_LINE "Can't read the KB statistics.");
      return nullptr;
    }
    std::vector<double> vB((size_t)T);
    if (fseeko(fc.f, (off_t)offB, SEEK_SET) != 0 || !rd(fc.f, vB.data(), vB.size() * 8) ||
        cudaMemcpy(E.dVB_, vB.data(), vB.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
      *err = FileOpErr(filePath, PQA_FILE_LINE "Can't read the _vB weights.");
      return nullptr;
    }
    if (fseeko(fc.f, (off_t)offTail, SEEK_SET) != 0) { *err = FileOpErr(filePath, PQA_FILE_LINE "Can't seek to the KB tail."); return nullptr; }
  }
  for (int g = 0; g < 2; g++) {   // question gaps, target gaps (BaseEngine::ReadGaps, BaseEngine.cpp:126-140)
    int64_t nGaps = 0;
    if (!rd(fc.f, &nGaps, 8)) { *err = FileOpErr(filePath, PQA_FILE_LINE "Can't read the gaps."); return nullptr; }
    const int64_t range = g == 0 ? Q : T;
    if (nGaps < 0 || nGaps > range) { *err = FileOpErr(filePath, PQA_FILE_LINE "Corrupt gap count."); return nullptr; }
    if (nGaps > 0 && eng->IsSharded()) { *err = ErrNotImplemented("B200 engine: a KB file with removed questions/targets (gaps) on a sharded engine; compact it first"); return nullptr; }
    std::vector<int64_t> ids((size_t)nGaps);
    if (nGaps > 0 && !rd(fc.f, ids.data(), ids.size() * 8)) { *err = FileOpErr(filePath, PQA_FILE_LINE "Can't read the gaps."); return nullptr; }
    GapSet &gs = g == 0 ? eng->qGaps_ : eng->tGaps_;
    for (int64_t id : ids) {
      if (id < 0 || id >= range || gs.IsGap(id)) { *err = FileOpErr(filePath, PQA_FILE_LINE "Corrupt gap list."); return nullptr; }
      gs.Release(id);
    }
  }
  if (!eng->pimQ_.Load(fc.f) || !eng->pimT_.Load(fc.f) || !eng->pimQuiz_.Load(fc.f)) {   // BaseEngine.cpp:45-56
    *err = FileOpErr(filePath, PQA_FILE_LINE "Can't read the permanent-compact ID mappings.");
    return nullptr;
  }
  eng->SyncGapBits();
  eng->nQuestionsAsked_.store(asked, std::memory_order_relaxed);
  return eng.release();
}

} // namespace pqa
