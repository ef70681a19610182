// probqa_b200: ShardGroup -- N shard engines in one process behind one engine handle (pqa_group.h).
// The group is an Engine shell: CheckQuiz / ValidateRecordAnswer / the call combiner / the quiz id registry are the base
// class's and run on the shell's own host registry, which evolves in lockstep with the registries of the shards because
// every shard is given exactly the same call sequence. Device work is fanned out through the shards' batch and P2P entry
// points (one host thread enqueues on all shards -- Begin -- and then collects -- End).
#include "pqa_group.h"

#include <algorithm>
#include <cstring>

namespace pqa {

#define SRC_LINE_STR2(x) #x
#define SRC_LINE_STR(x) SRC_LINE_STR2(x)
#define PQA_FILE_LINE "pqa_group.cu(" SRC_LINE_STR(__LINE__) "): "

#define PQA_TRY try {
#define PQA_CATCH_RETURN_ERR                                                              \
  } catch (const CudaFail &cf) { return ErrCuda(cf.code, cf.what(), cf.file, cf.line);    \
  } catch (const std::exception &ex) { return ErrStd(ex.what()); }

// CalcSplit (SRPoolRunner.h:96-110) over questions, or over 4-target vectors for target shards.
static void split_range(int64_t n, int64_t parts, int64_t p, int64_t *first, int64_t *count) {
  const int64_t quot = n / parts, rem = n % parts;
  *first = p * quot + std::min(p, rem);
  *count = quot + (p < rem ? 1 : 0);
}

std::vector<CiB200Options> ShardGroup::ShardOptions(const CiEngineDefinition &def, const CiB200Options &opts,
                                                    const CiB200GroupOptions &g) {
  if (g._nShards < 1 || g._nShards > kMaxPeers) throw std::runtime_error("probqa_b200: a shard group has 1.." + std::to_string(kMaxPeers) + " shards");
  if (g._axis != 0 && g._axis != 1) throw std::runtime_error("probqa_b200: shard axis must be 0 (questions) or 1 (targets)");
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev <= 0) throw std::runtime_error("probqa_b200: no CUDA device is visible; this engine has no CPU path");
  bool anyDevice = false;
  for (int r = 0; r < g._nShards; r++) anyDevice |= g._devices[r] >= 0;
  const int64_t units = g._axis == 0 ? def._nQuestions : (def._nTargets + 3) / 4;
  if (units < g._nShards) throw std::runtime_error("probqa_b200: more shards than " + std::string(g._axis == 0 ? "questions" : "4-target vectors"));
  std::vector<CiB200Options> out;
  for (int r = 0; r < g._nShards; r++) {
    CiB200Options o = opts;
    o._device = anyDevice ? g._devices[r] : (r % nDev);
    o._questionShardFirst = o._questionShardCount = o._targetShardFirst = o._targetShardCount = 0;
    int64_t first, count;
    split_range(units, g._nShards, r, &first, &count);
    if (g._axis == 0) { o._questionShardFirst = first; o._questionShardCount = count; }
    else { o._targetShardFirst = 4 * first; o._targetShardCount = std::min(4 * count, def._nTargets - 4 * first); }
    if (g._nShards == 1) o._questionShardCount = o._targetShardCount = 0;   // a group of one is a plain engine
    out.push_back(o);
  }
  return out;
}

ShardGroup::ShardGroup(const CiEngineDefinition &def, const CiB200Options &opts, const CiB200GroupOptions &gopts)
    : Engine(def, opts, ShellTag{}) {
  baseOpts_ = opts; groupOpts_ = gopts;
  const std::vector<CiB200Options> so = ShardOptions(def, opts, gopts);
  device_ = so[0]._device;
  for (const CiB200Options &o : so) shards_.emplace_back(new Engine(def, o));
  Connect(gopts);
}

ShardGroup *ShardGroup::LoadKBGroup(const char *filePath, const CiB200Options &opts, const CiB200GroupOptions &gopts, PqaError **err) {
  *err = nullptr;
  // the first shard reads the header; the shell and the other shards follow its dimensions
  CiB200Options probe = opts;
  FILE *f = filePath ? std::fopen(filePath, "rb") : nullptr;
  if (!f) { *err = MakeError(ErrCode::CantOpenFile, PQA_FILE_LINE "Can't open the KB file to read.", std::string("filePath=[") + (filePath ? filePath : "") + "]"); return nullptr; }
  uint64_t prec = 0, asked = 0;
  int64_t dims[3] = {0, 0, 0};
  const bool ok = std::fread(&prec, 8, 1, f) == 1 && std::fread(dims, 24, 1, f) == 1 && std::fread(&asked, 8, 1, f) == 1;
  std::fclose(f);
  if (!ok) { *err = MakeError(ErrCode::FileOp, PQA_FILE_LINE "Can't read the KB header.", std::string("filePath=[") + filePath + "]"); return nullptr; }
  if ((prec & 0xF) != 3) { *err = ErrNotImplemented("B200 engine on precision type other than double (KB file header)."); return nullptr; }
  if (dims[0] < 2 || dims[1] < 1 || dims[2] < 2) { *err = ErrInsufficientDims(dims[0], dims[1], dims[2]); return nullptr; }
  CiEngineDefinition def;
  std::memset(&def, 0, sizeof(def));
  def._nAnswers = dims[0]; def._nQuestions = dims[1]; def._nTargets = dims[2];
  def._precType = 3; def._precMantissa = (uint32_t)((prec >> 4) & 0xFFFFFFF); def._precExponent = (uint16_t)((prec >> 32) & 0xFFFF);
  def._initAmount = 1.0;
  std::unique_ptr<ShardGroup> g(new ShardGroup(def, probe));
  const std::vector<CiB200Options> so = ShardOptions(def, opts, gopts);
  g->device_ = so[0]._device;
  for (const CiB200Options &o : so) {
    Engine *e = Engine::LoadKB(filePath, o, err);
    if (!e) return nullptr;
    g->shards_.emplace_back(e);
  }
  g->nQuestionsAsked_.store(asked, std::memory_order_relaxed);
  g->baseOpts_ = opts; g->groupOpts_ = gopts;
  g->pimQ_ = g->shards_[0]->pimQ_; g->pimT_ = g->shards_[0]->pimT_; g->pimQuiz_ = g->shards_[0]->pimQuiz_;   // the id maps of the file
  g->Connect(gopts);
  return g.release();
}

ShardGroup::~ShardGroup() {
  for (auto &s : shards_) { cudaSetDevice(s->device()); delete s->Synchronize(); }
  shards_.clear();
}

// Inboxes of all shards, wired directly (one process: the device pointers are valid everywhere once peer access is on).
void ShardGroup::Connect(const CiB200GroupOptions &gopts) {
  cap_ = gopts._maxBatch > 0 ? gopts._maxBatch : 256;
  const int n = (int)shards_.size();
  if (n == 1) return;
  std::vector<void *> bases((size_t)n, nullptr);
  for (int r = 0; r < n; r++) {
    cudaSetDevice(shards_[r]->device());
    int64_t bytes = 0;
    if (PqaError *e = shards_[r]->P2PInit(r, n, cap_, &bases[(size_t)r], &bytes)) { const std::string m = e->ToString(true); delete e; throw std::runtime_error(m); }
  }
  for (int r = 0; r < n; r++)
    if (PqaError *e = (cudaSetDevice(shards_[r]->device()), shards_[r]->P2PConnect(bases.data()))) { const std::string m = e->ToString(true); delete e; throw std::runtime_error(m); }
  if (gopts._exactOrder && gopts._axis == 1)
    for (auto &s : shards_)
      if (PqaError *e = s->P2PSetExactOrder(1)) { const std::string m = e->ToString(true); delete e; throw std::runtime_error(m); }
}

// first error wins, the rest are released
static PqaError *NotMaintenanceGroup(const char *what) {
  return MakeError(ErrCode::WrongMode, std::string("Can't perform maintenance-only mode operation - ") + what +
                                           " - because current mode is not maintenance (but regular/shutdown?).");
}
static PqaError *Keep(PqaError *first, PqaError *next) {
  if (!first) return next;
  delete next;
  return first;
}

PqaError *ShardGroup::StartQuizBatch(int64_t n, int64_t *pQuizIds) {
  if (maintenance_) return WrongMode("Start/Resume quiz");
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds");
  std::lock_guard<std::mutex> lk(mu_);
  for (int64_t x = 0; x < n; x++) pQuizIds[x] = AssignQuizId();
  std::vector<int64_t> ids((size_t)n);
  PqaError *err = nullptr;
  for (auto &s : shards_) {
    cudaSetDevice(s->device());
    err = Keep(err, s->StartQuizBatch(n, ids.data()));
    if (!err && !std::equal(ids.begin(), ids.end(), pQuizIds))
      err = MakeError(ErrCode::Internal, PQA_FILE_LINE "the shards' quiz registries have diverged");
  }
  return err;
}

// ResumeQuiz (CpuEngine.cpp:277-282, CECreateQuizOperation.cpp:55-83): the likelihood product over the answered questions is
// elementwise per target, but its cells are spread over the shards (rows of some questions, or columns of all). Shard 0
// runs the single-engine kernel once with a cell view over ALL shards (peer pointers; the devices of a group have peer
// access) and stores the finished rows into every shard's replica of the quiz; the other shards only keep their
// registries in step. Bit-identical to one engine by construction: it is the same kernel on the same cells.
PqaError *ShardGroup::ResumeQuizBatch(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, int64_t *pQuizIds) {
  if (maintenance_) return WrongMode("Start/Resume quiz");
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pCounts || !pQuizIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pCounts/pQuizIds");
  if (shards_.size() == 1) {
    std::lock_guard<std::mutex> lk(mu_);
    PqaError *e = (cudaSetDevice(shards_[0]->device()), shards_[0]->ResumeQuizBatchEx(n, pCounts, pAQs, pQuizIds, nullptr, nullptr, nullptr));
    MirrorResumed(n, pCounts, pAQs, pQuizIds);
    return e;
  }
  std::lock_guard<std::mutex> lk(mu_);
  ResumeSource src;
  PoolList pools;
  src.nShards = pools.n = (int)shards_.size();
  src.K = K_;
  for (size_t r = 0; r < shards_.size(); r++) {
    const DeviceKB k = shards_[r]->kb();
    src.sA[r] = k.sA; src.mD[r] = k.mD; src.qFirst[r] = k.qFirst; src.qCount[r] = k.qCount;
    src.tFirst[r] = shards_[r]->targetShardFirst(); src.TpL[r] = k.Tp;
    pools.p[r] = shards_[r]->pool();
  }
  // every shard grows its quiz pool first: the leader's kernel writes into all of them
  std::vector<int> status;
  std::vector<int64_t> ids((size_t)n, -1);
  cudaSetDevice(shards_[0]->device());
  for (size_t r = 1; r < shards_.size(); r++) {     // pools of the followers must hold the new slots before the kernel runs
    cudaSetDevice(shards_[r]->device());
    std::lock_guard<std::mutex> lkS(shards_[r]->mu_);
    shards_[r]->EnsureQuizCapacity((int64_t)shards_[r]->quizzes_.size() + n);
    pools.p[r] = shards_[r]->pool();
  }
  cudaSetDevice(shards_[0]->device());
  {
    std::lock_guard<std::mutex> lkS(shards_[0]->mu_);
    shards_[0]->EnsureQuizCapacity((int64_t)shards_[0]->quizzes_.size() + n);
    pools.p[0] = shards_[0]->pool();
  }
  PqaError *err = shards_[0]->ResumeQuizBatchEx(n, pCounts, pAQs, pQuizIds, &src, &pools, &status);
  if (err && err->code != ErrCode::I64Underflow) return err;      // validation failure: nothing was assigned anywhere
  for (size_t r = 1; r < shards_.size(); r++) {
    cudaSetDevice(shards_[r]->device());
    std::vector<int> st = status;
    PqaError *e = shards_[r]->ResumeQuizBatchEx(n, pCounts, pAQs, ids.data(), nullptr, nullptr, &st);
    if (e && e->code == ErrCode::I64Underflow) { delete e; e = nullptr; }       // reported once, by the leader
    if (!e && !std::equal(ids.begin(), ids.end(), pQuizIds)) e = MakeError(ErrCode::Internal, PQA_FILE_LINE "the shards' quiz registries have diverged");
    if (e) { delete err; return e; }
  }
  MirrorResumed(n, pCounts, pAQs, pQuizIds);
  return err;
}

// the shell's own registry follows the shards': same ids in the same order (failed quizzes are assigned and released again)
void ShardGroup::MirrorResumed(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, const int64_t *pQuizIds) {
  std::vector<int64_t> mine((size_t)n);
  for (int64_t x = 0; x < n; x++) mine[(size_t)x] = AssignQuizId();
  int64_t off = 0;
  for (int64_t x = 0; x < n; x++) {
    HostQuiz &q = quizzes_[(size_t)mine[(size_t)x]];
    for (int64_t y = 0; y < pCounts[x]; y++) q.answers.push_back(pAQs[off + y]);
    off += pCounts[x];
  }
  for (int64_t x = 0; x < n; x++) {
    if (pQuizIds[x] >= 0) continue;                    // the reference's I64Underflow: the quiz was unassigned again
    HostQuiz &q = quizzes_[(size_t)mine[(size_t)x]];
    q.present = false; q.answers.clear(); q.activeQuestion = -1;
    quizGaps_.push_back(mine[(size_t)x]);
    pimQuiz_.RemoveComp(mine[(size_t)x]);
  }
}

// BaseEngine::ClearOldQuizzes (BaseEngine.cpp:814-872) is registry work: the shell decides with its own usage times which
// quizzes go (the base class's algorithm), and the shards release exactly those, in the same order.
PqaError *ShardGroup::ClearOldQuizzes(int64_t maxCount, double maxAgeSec) {
  size_t before;
  { std::lock_guard<std::mutex> lk(mu_); before = quizGaps_.size(); }
  if (PqaError *e = Engine::ClearOldQuizzes(maxCount, maxAgeSec)) return e;
  std::vector<int64_t> released;
  { std::lock_guard<std::mutex> lk(mu_); released.assign(quizGaps_.begin() + (std::ptrdiff_t)before, quizGaps_.end()); }
  PqaError *err = nullptr;
  if (!released.empty())
    for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->ReleaseQuizBatch((int64_t)released.size(), released.data())); }
  return err;
}

PqaError *ShardGroup::NextQuestionBatch(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms, int64_t *pQuestions,
                                        void **ppErrors) {
  if (maintenance_) return WrongMode("compute next question");
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pQuestions) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pQuestions");
  std::lock_guard<std::mutex> lk(mu_);
  PQA_TRY
  std::vector<int64_t> valid, where;
  std::vector<uint64_t> rnd;
  PqaError *firstErr = nullptr;
  for (int64_t x = 0; x < n; x++) {            // quizzes that fail validation get their own error and are left out
    pQuestions[x] = -1;
    if (ppErrors) ppErrors[x] = nullptr;
    if (PqaError *e = CheckQuiz(pQuizIds[x])) {
      if (ppErrors) ppErrors[x] = e; else firstErr = Keep(firstErr, e);
      continue;
    }
    valid.push_back(pQuizIds[x]); where.push_back(x);
    rnd.push_back(pRandoms ? pRandoms[x] : NextRandom());
  }
  if (broken_) return Broken();
  uint64_t nAsked = 0;
  for (size_t s0 = 0; s0 < valid.size(); s0 += (size_t)cap_) {      // slices of the inbox capacity
    const int64_t m = (int64_t)std::min(valid.size() - s0, (size_t)cap_);
    std::vector<int64_t> out((size_t)m, -1);
    std::vector<void *> errs((size_t)m, nullptr);
    PqaError *sliceErr = nullptr;
    if (shards_.size() == 1) {
      sliceErr = (cudaSetDevice(shards_[0]->device()), shards_[0]->NextQuestionBatch(m, valid.data() + s0, rnd.data() + s0, out.data(), errs.data()));
    } else {
      for (auto &s : shards_) { cudaSetDevice(s->device()); sliceErr = Keep(sliceErr, s->P2PNextQuestionBegin(m, valid.data() + s0, rnd.data() + s0)); }
      for (size_t r = 0; r < shards_.size(); r++) {
        cudaSetDevice(shards_[r]->device());
        std::vector<int64_t> o2((size_t)m, -1);
        PqaError *e = shards_[r]->P2PNextQuestionEnd(m, valid.data() + s0, r == 0 ? out.data() : o2.data(), r == 0 ? errs.data() : nullptr);
        if (r != 0 && e && e->code == ErrCode::QuestionsExhausted) { delete e; e = nullptr; }   // reported per quiz by shard 0
        if (!e && r != 0 && o2 != out) e = MakeError(ErrCode::Internal, PQA_FILE_LINE "the shards selected different questions");
        sliceErr = Keep(sliceErr, e);
      }
      if (sliceErr && sliceErr->code != ErrCode::QuestionsExhausted) broken_ = true;   // a shard failed mid-exchange: no lockstep any more
    }
    for (int64_t x = 0; x < m; x++) {
      const int64_t at = where[s0 + (size_t)x];
      if (errs[(size_t)x]) {
        if (ppErrors) ppErrors[at] = errs[(size_t)x]; else firstErr = Keep(firstErr, static_cast<PqaError *>(errs[(size_t)x]));
        continue;
      }
      if (sliceErr || out[(size_t)x] < 0) continue;
      pQuestions[at] = out[(size_t)x];
      quizzes_[(size_t)valid[s0 + (size_t)x]].activeQuestion = out[(size_t)x];
      nAsked++;
    }
    firstErr = Keep(firstErr, sliceErr);
  }
  nQuestionsAsked_.fetch_add(nAsked, std::memory_order_relaxed);
  return firstErr;
  PQA_CATCH_RETURN_ERR
}

PqaError *ShardGroup::RecordAnswerBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) {
  if (maintenance_) return WrongMode("record an answer");
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pAnswers) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pAnswers");
  std::lock_guard<std::mutex> lk(mu_);
  PQA_TRY
  if (PqaError *e = ValidateRecordAnswer(n, pQuizIds, pAnswers)) return e;
  if (broken_) return Broken();
  PqaError *err = nullptr;
  for (int64_t s0 = 0; s0 < n && !err; s0 += cap_) {
    const int64_t m = std::min(cap_, n - s0);
    if (shards_.size() == 1) {
      err = (cudaSetDevice(shards_[0]->device()), shards_[0]->RecordAnswerBatch(m, pQuizIds + s0, pAnswers + s0));
    } else {
      for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->P2PRecordAnswerBegin(m, pQuizIds + s0, pAnswers + s0)); }
      for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->P2PRecordAnswerEnd()); }
      if (err) broken_ = true;       // the shards' posteriors and registries may differ now
    }
  }
  if (err) return err;
  for (int64_t x = 0; x < n; x++) {
    HostQuiz &q = quizzes_[(size_t)pQuizIds[x]];
    q.answers.push_back(CiAnsweredQuestion{q.activeQuestion, pAnswers[x]});
    q.activeQuestion = -1;
  }
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *ShardGroup::SetActiveQuestionBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pQuestions) {
  if (maintenance_) return WrongMode("set active question");
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pQuestions) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pQuestions");
  std::lock_guard<std::mutex> lk(mu_);
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->SetActiveQuestionBatch(n, pQuizIds, pQuestions)); }
  if (!err)
    for (int64_t x = 0; x < n; x++) quizzes_[(size_t)pQuizIds[x]].activeQuestion = pQuestions[x];
  return err;
}

PqaError *ShardGroup::ListTopTargetsBatch(int64_t n, const int64_t *pQuizIds, int64_t maxCount, CiRatedTarget *pDest,
                                          int64_t *pCounts) {
  if (maintenance_) return WrongMode("compute next question");
  if (n > 0 && pQuizIds) {
    std::lock_guard<std::mutex> lk(mu_);
    for (int64_t x = 0; x < n; x++)
      if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  }
  return (cudaSetDevice(shards_[0]->device()), shards_[0]->ListTopTargetsBatch(n, pQuizIds, maxCount, pDest, pCounts));   // quiz state is replicated: any shard answers
}

PqaError *ShardGroup::RecordQuizTargetBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pTargets, const double *pAmounts) {
  if (maintenance_) return WrongMode("record quiz target");
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n == 0) return nullptr;
  if (!pQuizIds || !pTargets) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds/pTargets");
  std::lock_guard<std::mutex> lk(mu_);
  for (int64_t x = 0; x < n; x++) {
    if (pAmounts && !(pAmounts[x] > 0)) return ErrNonPositiveAmount(pAmounts[x], PQA_FILE_LINE "|amount| must be positive.");
    if (pTargets[x] < 0 || pTargets[x] >= T_) return ErrIndexOutOfRange(pTargets[x], 0, T_ - 1, PQA_FILE_LINE "Target index is not in KB range.");
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  }
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->RecordQuizTargetBatch(n, pQuizIds, pTargets, pAmounts)); }   // each shard: its own cells, vB everywhere
  return err;
}

PqaError *ShardGroup::Train(int64_t nQuestions, const CiAnsweredQuestion *pAQs, int64_t iTarget, double amount) {
  if (maintenance_) return WrongMode("train");
  // every shard validates the same arguments the same way: if the first one refuses, nothing was applied anywhere
  cudaSetDevice(shards_[0]->device());
  if (PqaError *e = shards_[0]->Train(nQuestions, pAQs, iTarget, amount)) return e;
  PqaError *err = nullptr;
  for (size_t r = 1; r < shards_.size(); r++) { cudaSetDevice(shards_[r]->device()); err = Keep(err, shards_[r]->Train(nQuestions, pAQs, iTarget, amount)); }
  if (!err) nQuestionsAsked_.fetch_add((uint64_t)std::max<int64_t>(nQuestions, 0), std::memory_order_relaxed);   // CpuEngine.cpp:179
  return err;
}

PqaError *ShardGroup::ReleaseQuizBatch(int64_t n, const int64_t *pQuizIds) {
  if (maintenance_) return WrongMode("release quiz");
  if (n < 0) return ErrNegativeCount(n, PQA_FILE_LINE "|n| must be non-negative.");
  if (n > 0 && !pQuizIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQuizIds");
  std::lock_guard<std::mutex> lk(mu_);
  for (int64_t x = 0; x < n; x++)
    if (PqaError *e = CheckQuiz(pQuizIds[x])) return e;
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->ReleaseQuizBatch(n, pQuizIds)); }
  if (err) return err;
  for (int64_t x = 0; x < n; x++) {
    HostQuiz &q = quizzes_[(size_t)pQuizIds[x]];
    q.present = false; q.answers.clear(); q.answers.shrink_to_fit(); q.activeQuestion = -1;
    quizGaps_.push_back(pQuizIds[x]);
    pimQuiz_.RemoveComp(pQuizIds[x]);
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------------------
// Maintenance mode on a group (pqa_group.h).
void ShardGroup::MoveCells(Engine &whole, bool toWhole) {
  for (auto &sp : shards_) {
    Engine &s = *sp;
    cudaSetDevice(s.device());
    const int64_t rowsA = s.qLocal_ * K_, rowsD = s.qLocal_;
    double *wA = whole.dSA_ + (s.qFirst_ * K_) * whole.TpL_ + s.tFirst_, *wD = whole.dMD_ + s.qFirst_ * whole.TpL_ + s.tFirst_;
    const size_t wPitch = (size_t)whole.TpL_ * 8, sPitch = (size_t)s.TpL_ * 8, width = (size_t)s.tLocal_ * 8;
    if (toWhole) {
      PQA_CU(cudaMemcpy2D(wA, wPitch, s.dSA_, sPitch, width, (size_t)rowsA, cudaMemcpyDefault));
      PQA_CU(cudaMemcpy2D(wD, wPitch, s.dMD_, sPitch, width, (size_t)rowsD, cudaMemcpyDefault));
    } else {
      PQA_CU(cudaMemcpy2D(s.dSA_, sPitch, wA, wPitch, width, (size_t)rowsA, cudaMemcpyDefault));
      PQA_CU(cudaMemcpy2D(s.dMD_, sPitch, wD, wPitch, width, (size_t)rowsD, cudaMemcpyDefault));
      PQA_CU(cudaMemcpy(s.dVB_, whole.dVB_, sizeof(double) * (size_t)whole.Tp_, cudaMemcpyDefault));
      s.MarkKBChanged();
    }
  }
  if (toWhole) {
    cudaSetDevice(whole.device());
    PQA_CU(cudaMemcpy(whole.dVB_, shards_[0]->dVB_, sizeof(double) * (size_t)whole.Tp_, cudaMemcpyDefault));
    whole.MarkKBChanged();
  }
}

void ShardGroup::MirrorMaintenanceState() {
  Q_ = maint_->Q_; T_ = maint_->T_; Tp_ = maint_->Tp_; TpL_ = Tp_; qLocal_ = Q_; tLocal_ = T_;
  askedWords_ = maint_->askedWords_;
  qGaps_ = maint_->qGaps_; tGaps_ = maint_->tGaps_;
  pimQ_ = maint_->pimQ_; pimT_ = maint_->pimT_;
}

PqaError *ShardGroup::StartMaintenance(bool forceQuizzes) {
  if (broken_) return Broken();
  PQA_TRY
  if (PqaError *e = Engine::StartMaintenance(forceQuizzes)) return e;     // the shell: mode switch, quizzes (BaseEngine.cpp:640-683)
  CiEngineDefinition def;
  std::memset(&def, 0, sizeof(def));
  def._nAnswers = K_; def._nQuestions = Q_; def._nTargets = T_; def._precType = 3;
  def._precMantissa = precMantissa_; def._precExponent = precExponent_; def._initAmount = initAmount_;
  CiB200Options o = baseOpts_;
  o._device = shards_[0]->device(); o._emulatedWorkers = W_;
  o._questionShardFirst = o._questionShardCount = o._targetShardFirst = o._targetShardCount = 0;
  cudaSetDevice(o._device);
  maint_.reset(new Engine(def, o));
  maint_->pimQ_ = pimQ_; maint_->pimT_ = pimT_; maint_->pimQuiz_ = pimQuiz_; maint_->qGaps_ = qGaps_; maint_->tGaps_ = tGaps_;
  maint_->nQuestionsAsked_.store(nQuestionsAsked_.load(std::memory_order_relaxed), std::memory_order_relaxed);
  MoveCells(*maint_, true);
  shards_.clear();                                                       // their device memory is free for the resized KB
  if (PqaError *e = maint_->StartMaintenance(true)) return e;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *ShardGroup::FinishMaintenance() {
  if (!maintenance_ || !maint_) return MakeError(ErrCode::MaintenanceModeAlreadyThis, PQA_FILE_LINE "The engine is in regular mode already.", "activeMode=0");
  if (maint_->qGaps_.GetNGaps() > 0 || maint_->tGaps_.GetNGaps() > 0)
    return ErrNotImplemented("sharded engine group: removed questions / targets (gaps) on shards; call Compact before FinishMaintenance");
  PQA_TRY
  if (PqaError *e = maint_->FinishMaintenance()) return e;
  MirrorMaintenanceState();
  CiEngineDefinition def;
  std::memset(&def, 0, sizeof(def));
  def._nAnswers = K_; def._nQuestions = Q_; def._nTargets = T_; def._precType = 3;
  def._precMantissa = precMantissa_; def._precExponent = precExponent_; def._initAmount = initAmount_;
  CiB200Options o = baseOpts_;
  o._emulatedWorkers = W_;
  const std::vector<CiB200Options> so = ShardOptions(def, o, groupOpts_);
  for (const CiB200Options &x : so) { cudaSetDevice(x._device); shards_.emplace_back(new Engine(def, x)); }
  MoveCells(*maint_, false);
  for (auto &s : shards_) { s->pimQ_ = pimQ_; s->pimT_ = pimT_; s->pimQuiz_ = pimQuiz_; s->nQuestionsAsked_.store(nQuestionsAsked_.load(std::memory_order_relaxed), std::memory_order_relaxed); }
  maint_.reset();
  Connect(groupOpts_);
  maintenance_ = false;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

PqaError *ShardGroup::AddQsTs(int64_t nQuestions, CiAddQorTParam *pAqps, int64_t nTargets, CiAddQorTParam *pAtps) {
  if (!maintenance_ || !maint_) return NotMaintenanceGroup("add questions/targets");
  PqaError *e = (cudaSetDevice(maint_->device()), maint_->AddQsTs(nQuestions, pAqps, nTargets, pAtps));
  MirrorMaintenanceState();
  return e;
}
PqaError *ShardGroup::RemoveQuestions(int64_t nQuestions, const int64_t *pQIds) {
  if (!maintenance_ || !maint_) return NotMaintenanceGroup("remove questions");
  PqaError *e = (cudaSetDevice(maint_->device()), maint_->RemoveQuestions(nQuestions, pQIds));
  MirrorMaintenanceState();
  return e;
}
PqaError *ShardGroup::RemoveTargets(int64_t nTargets, const int64_t *pTIds) {
  if (!maintenance_ || !maint_) return NotMaintenanceGroup("remove targets");
  PqaError *e = (cudaSetDevice(maint_->device()), maint_->RemoveTargets(nTargets, pTIds));
  MirrorMaintenanceState();
  return e;
}
PqaError *ShardGroup::Compact(int64_t *pnQuestions, const int64_t **ppOldQuestions, int64_t *pnTargets, const int64_t **ppOldTargets) {
  if (!maintenance_ || !maint_) return NotMaintenanceGroup("compact the KB");
  PqaError *e = (cudaSetDevice(maint_->device()), maint_->Compact(pnQuestions, ppOldQuestions, pnTargets, ppOldTargets));
  MirrorMaintenanceState();
  return e;
}

// KB access: whole-KB host arrays, every shard fills / takes its own rows or columns of them
PqaError *ShardGroup::UploadKB(const double *sA, const double *mD, const double *vB) {
  if (maint_) return (cudaSetDevice(maint_->device()), maint_->UploadKB(sA, mD, vB));
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->UploadKB(sA, mD, vB)); }
  return err;
}
PqaError *ShardGroup::DownloadKB(double *sA, double *mD, double *vB) {
  if (maint_) return (cudaSetDevice(maint_->device()), maint_->DownloadKB(sA, mD, vB));
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->DownloadKB(sA, mD, vB)); }
  return err;
}
PqaError *ShardGroup::CopyATargets(int64_t iQuestion, int64_t iAnswer, int64_t maxTargets, double *pFreqs) {
  if (iQuestion < 0 || iQuestion >= Q_) return ErrIndexOutOfRange(iQuestion, 0, Q_ - 1, PQA_FILE_LINE "Question index is not in KB range.");
  if (iAnswer < 0 || iAnswer >= K_) return ErrIndexOutOfRange(iAnswer, 0, K_ - 1, PQA_FILE_LINE "Answer index is not in KB range.");
  if (maint_) return (cudaSetDevice(maint_->device()), maint_->CopyATargets(iQuestion, iAnswer, maxTargets, pFreqs));
  PqaError *err = nullptr;
  for (auto &s : shards_) {
    cudaSetDevice(s->device());
    if (s->questionShardCount() != Q_ && (iQuestion < s->questionShardFirst() || iQuestion >= s->questionShardFirst() + s->questionShardCount())) continue;
    err = Keep(err, s->CopyATargets(iQuestion, iAnswer, maxTargets, pFreqs));
  }
  return err;
}
PqaError *ShardGroup::CopyDTargets(int64_t iQuestion, int64_t maxTargets, double *pFreqs) {
  if (iQuestion < 0 || iQuestion >= Q_) return ErrIndexOutOfRange(iQuestion, 0, Q_ - 1, PQA_FILE_LINE "Question index is not in KB range.");
  if (maint_) return (cudaSetDevice(maint_->device()), maint_->CopyDTargets(iQuestion, maxTargets, pFreqs));
  PqaError *err = nullptr;
  for (auto &s : shards_) {
    cudaSetDevice(s->device());
    if (s->questionShardCount() != Q_ && (iQuestion < s->questionShardFirst() || iQuestion >= s->questionShardFirst() + s->questionShardCount())) continue;
    err = Keep(err, s->CopyDTargets(iQuestion, maxTargets, pFreqs));
  }
  return err;
}
PqaError *ShardGroup::CopyBTargets(int64_t maxTargets, double *pFreqs) {
  Engine *e = maint_ ? maint_.get() : shards_[0].get();
  return (cudaSetDevice(e->device()), e->CopyBTargets(maxTargets, pFreqs));
}
PqaError *ShardGroup::CopyQuizPriors(int64_t iQuiz, double *pPriors) { if (maint_) return WrongMode("copy quiz priors"); return (cudaSetDevice(shards_[0]->device()), shards_[0]->CopyQuizPriors(iQuiz, pPriors)); }

PqaError *ShardGroup::SaveKB(const char *filePath) {
  if (maint_) return (cudaSetDevice(maint_->device()), maint_->SaveKB(filePath));
  if (shards_.size() == 1) return (cudaSetDevice(shards_[0]->device()), shards_[0]->SaveKB(filePath));
  PqaError *err = (cudaSetDevice(shards_[0]->device()), shards_[0]->SaveKBShard(filePath, true));       // frame + its cells, then the others in place
  for (size_t r = 1; r < shards_.size() && !err; r++) { cudaSetDevice(shards_[r]->device()); err = shards_[r]->SaveKBShard(filePath, false); }
  return err;
}
PqaError *ShardGroup::Shutdown(const char *saveFilePath) {
  if (shutdown_.exchange(true, std::memory_order_acq_rel))
    return MakeError(ErrCode::ObjectShutDown, std::string("MaintenanceSwitch seems already shut down.") +
                     (saveFilePath ? std::string(" Not saving file: ") + saveFilePath : std::string()), "CpuEngine<taNumber>::Shutdown()");
  PqaError *err = nullptr;
  if (saveFilePath && *saveFilePath) err = SaveKB(saveFilePath);
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->Shutdown(nullptr)); }
  std::lock_guard<std::mutex> lk(mu_);
  quizzes_.clear(); quizGaps_.clear();
  pimQuiz_.OnCompact(0, nullptr);
  return err;
}
PqaError *ShardGroup::FillBinarySearchKB(double rounds) {
  if (maint_) return (cudaSetDevice(maint_->device()), maint_->FillBinarySearchKB(rounds));
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->FillBinarySearchKB(rounds)); }
  return err;
}
PqaError *ShardGroup::Synchronize() {
  if (maint_) return (cudaSetDevice(maint_->device()), maint_->Synchronize());
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->Synchronize()); }
  return err;
}
PqaError *ShardGroup::SetEvalKernel(int32_t which, int64_t chunkTargets, int64_t quizzesPerCta, int32_t kahanLanesPerThread) {
  PqaError *err = nullptr;
  for (auto &s : shards_) { cudaSetDevice(s->device()); err = Keep(err, s->SetEvalKernel(which, chunkTargets, quizzesPerCta, kahanLanesPerThread)); }
  return err;
}

// every shard runs the same selection on the same priorities: the first shard's counters are the group's
PqaError *ShardGroup::AnomalyCounts(uint64_t *pCounts3) {
  if (maint_) return maint_->AnomalyCounts(pCounts3);
  if (shards_.empty()) { for (int a = 0; a < kAnomalyKinds; a++) pCounts3[a] = 0; return nullptr; }
  cudaSetDevice(shards_[0]->device());
  return shards_[0]->AnomalyCounts(pCounts3);
}

} // namespace pqa
