// probqa_b200: maintenance mode of the B200 engine -- KB resize (AddQsTs), removal of questions / targets (gaps),
// compaction, and the permanent <-> compact id maps. Mirrors BaseEngine (PqaCore/BaseEngine.cpp:154-218, 640-779),
// CpuEngine::AddQsTsSpec / CompactSpec (PqaCore/CpuEngine.cpp:468-658), GapTracker (GapTracker.h) and
// PermanentIdManager (PermanentIdManager.cpp). The KB cells live on the device: resizing allocates new sA/mD/vB arrays
// with the new row stride and copies / fills them with kernels; compaction is one gather per array. Removed targets and
// questions are "gaps": bitmaps the quiz kernels mask with (pqa_kernels.cuh DeviceKB::tgaps / qgaps), exactly like the
// reference's gap masks. Citations are relative to /root/reference/ProbQA/.
#include "pqa_engine.h"

#include <algorithm>
#include <cstring>
#include <ctime>
#include <memory>

namespace pqa {

#define SRC_LINE_STR2(x) #x
#define SRC_LINE_STR(x) SRC_LINE_STR2(x)
#define PQA_FILE_LINE "pqa_maint.cu(" SRC_LINE_STR(__LINE__) "): "

#define PQA_TRY try {
#define PQA_CATCH_RETURN_ERR                                                              \
  } catch (const CudaFail &cf) { return ErrCuda(cf.code, cf.what(), cf.file, cf.line);    \
  } catch (const std::exception &ex) { return ErrStd(ex.what()); }

// ---------------------------------------------------------------------------------------------------------
// PermanentIdManager (PermanentIdManager.cpp)
int64_t PermIds::PermFromComp(int64_t comp) const {                       // :12-17
  if (comp < 0 || comp >= (int64_t)comp2perm_.size()) return -1;
  return comp2perm_[(size_t)comp];
}
int64_t PermIds::CompFromPerm(int64_t perm) const {                       // :19-25
  auto it = perm2comp_.find(perm);
  return it == perm2comp_.end() ? -1 : it->second;
}
bool PermIds::RemoveComp(int64_t comp) {                                  // :71-90
  if (comp < 0 || comp >= (int64_t)comp2perm_.size()) return false;
  const int64_t perm = comp2perm_[(size_t)comp];
  if (perm == -1) return false;
  auto it = perm2comp_.find(perm);
  if (it == perm2comp_.end()) return false;
  perm2comp_.erase(it);
  comp2perm_[(size_t)comp] = -1;
  return true;
}
bool PermIds::RenewComp(int64_t comp) {                                   // :92-106
  if (comp < 0 || comp >= (int64_t)comp2perm_.size()) return false;
  if (comp2perm_[(size_t)comp] != -1) return false;
  comp2perm_[(size_t)comp] = nextPerm_;
  perm2comp_.emplace(nextPerm_, comp);
  nextPerm_++;
  return true;
}
bool PermIds::GrowTo(int64_t nComp) {                                     // :108-119
  if (nComp < (int64_t)comp2perm_.size()) return false;
  for (int64_t i = (int64_t)comp2perm_.size(); i < nComp; i++) {
    comp2perm_.push_back(nextPerm_);
    perm2comp_.emplace(nextPerm_, i);
    nextPerm_++;
  }
  return true;
}
bool PermIds::OnCompact(int64_t nNew, const int64_t *pOldIds) {           // :121-165
  if (nNew > (int64_t)comp2perm_.size() || nNew != (int64_t)perm2comp_.size()) return false;
  for (int64_t i = 0; i < nNew; i++) {
    const int64_t oldComp = pOldIds[i];
    if (oldComp < 0 || oldComp >= (int64_t)comp2perm_.size()) return false;
    const int64_t oldPerm = comp2perm_[(size_t)oldComp];
    if (oldPerm == -1) return false;
    auto it = perm2comp_.find(oldPerm);
    if (it == perm2comp_.end()) return false;
    it->second = i;
  }
  comp2perm_.assign((size_t)nNew, -1);
  for (const auto &m : perm2comp_) {
    if (m.second < 0 || m.second >= nNew) return false;
    comp2perm_[(size_t)m.second] = m.first;
  }
  return true;
}
bool PermIds::EnsurePermIdGreater(int64_t bound) {                        // :63-69
  if (nextPerm_ <= bound) { nextPerm_ = bound + 1; return true; }
  return false;
}
bool PermIds::RemapPermId(int64_t srcPerm, int64_t destPerm) {            // :167-184
  if (destPerm >= nextPerm_) return false;
  if (perm2comp_.find(destPerm) != perm2comp_.end()) return false;
  auto it = perm2comp_.find(srcPerm);
  if (it == perm2comp_.end()) return false;
  const int64_t comp = it->second;
  perm2comp_.erase(it);
  perm2comp_.emplace(destPerm, comp);
  comp2perm_[(size_t)comp] = destPerm;
  return true;
}
bool PermIds::Save(FILE *f, bool empty) const {                           // :27-39
  const int64_t nComp = empty ? 0 : (int64_t)comp2perm_.size();
  if (std::fwrite(&nextPerm_, 8, 1, f) != 1 || std::fwrite(&nComp, 8, 1, f) != 1) return false;
  return nComp == 0 || (int64_t)std::fwrite(comp2perm_.data(), 8, (size_t)nComp, f) == nComp;
}
bool PermIds::Load(FILE *f) {                                             // :41-61
  int64_t nComp = 0;
  if (std::fread(&nextPerm_, 8, 1, f) != 1 || std::fread(&nComp, 8, 1, f) != 1 || nComp < 0) return false;
  comp2perm_.resize((size_t)nComp);
  perm2comp_.clear();
  if (nComp > 0 && (int64_t)std::fread(comp2perm_.data(), 8, (size_t)nComp, f) != nComp) return false;
  for (int64_t i = 0; i < nComp; i++)
    if (comp2perm_[(size_t)i] != -1) perm2comp_.emplace(comp2perm_[(size_t)i], i);
  return true;
}

// ---------------------------------------------------------------------------------------------------------
PqaError *Engine::WrongMode(const char *what) const {
  return MakeError(ErrCode::WrongMode, std::string("Can't perform regular-only mode operation (") + what +
                                           ") because current mode is not regular (but maintenance/shutdown?).");
}

static PqaError *NotMaintenance(const char *what) {
  return MakeError(ErrCode::WrongMode, std::string("Can't perform maintenance-only mode operation - ") + what +
                                           " - because current mode is not maintenance (but regular/shutdown?).");
}

bool Engine::MapIds(int kind, bool permFromComp, int64_t count, int64_t *pIds) {   // BaseEngine.cpp:154-206
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  const PermIds &pim = kind == 0 ? pimQ_ : kind == 1 ? pimT_ : pimQuiz_;
  for (int64_t i = 0; i < count; i++) pIds[i] = permFromComp ? pim.PermFromComp(pIds[i]) : pim.CompFromPerm(pIds[i]);
  return true;
}
bool Engine::EnsurePermQuizGreater(int64_t bound) {                                // :208-212
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  return pimQuiz_.EnsurePermIdGreater(bound);
}
bool Engine::RemapQuizPermId(int64_t srcPermId, int64_t destPermId) {              // :214-218
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  return pimQuiz_.RemapPermId(srcPermId, destPermId);
}

// Bit j of word j>>5 = target j is a gap; same for questions. Uploaded only while gaps exist (kb() hands the kernels a
// null bitmap otherwise, which costs them nothing).
void Engine::SyncGapBits() {
  MarkKBChanged();   // gap targets are +0 lanes of the derived KB; every maintenance change of sA / mD ends here too
  auto upload = [&](const GapSet &g, int64_t n, DevBuf<uint32_t> &buf) {
    if (g.GetNGaps() == 0) return;
    std::vector<uint32_t> words((size_t)((n + 31) >> 5) + 1, 0u);
    for (int64_t i = 0; i < n; i++)
      if (g.IsGap(i)) words[(size_t)(i >> 5)] |= 1u << (i & 31);
    buf.ensure(words.size(), stream_);
    PQA_CU(cudaMemcpyAsync(buf.get(), words.data(), words.size() * 4, cudaMemcpyHostToDevice, stream_));
    PQA_CU(cudaStreamSynchronize(stream_));
  };
  upload(qGaps_, Q_, dQGapBits_);
  upload(tGaps_, T_, dTGapBits_);
}

void Engine::DropQuizPool() {
  if (quizCap_ > 0) {
    PQA_CU(cudaStreamSynchronize(stream_));
    cudaFree(dPriors_); cudaFree(dLogPriors_); cudaFree(dAsked_); cudaFree(dActive_); cudaFree(dNormS_);
    dPriors_ = dLogPriors_ = dNormS_ = nullptr; dAsked_ = nullptr; dActive_ = nullptr;
    quizCap_ = 0;
  }
  residentN_ = 0;
}

// BaseEngine::ClearOldQuizzes (BaseEngine.cpp:814-872): quizzes idle for more than maxAgeSec are released; if more than
// maxCount remain, the oldest are released until maxCount are left. A no-op outside regular mode.
PqaError *Engine::ClearOldQuizzes(int64_t maxCount, double maxAgeSec) {
  if (maxCount < 0) return ErrNegativeCount(maxCount, PQA_FILE_LINE "The number of quizzes to keep cannot be less than 0.");
  if (maintenance_) return nullptr;
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  struct QuizAge {
    int64_t iQuiz; double ageSec;
    bool operator<(const QuizAge &o) const { return ageSec < o.ageSec; }
  };
  std::vector<QuizAge> ages;
  const time_t callTime = std::time(nullptr);
  auto release = [&](int64_t i) {                          // LockedReleaseQuiz, :803-812
    HostQuiz &q = quizzes_[(size_t)i];
    q.present = false; q.answers.clear(); q.answers.shrink_to_fit(); q.activeQuestion = -1;
    quizGaps_.push_back(i);
    pimQuiz_.RemoveComp(i);
  };
  for (int64_t i = 0; i < (int64_t)quizzes_.size(); i++) {
    if (!quizzes_[(size_t)i].present) continue;
    const double ageSec = std::difftime(callTime, quizzes_[(size_t)i].lastUsage);
    if (ageSec > maxAgeSec) { release(i); continue; }
    ages.push_back(QuizAge{i, ageSec});
  }
  if ((int64_t)ages.size() > maxCount) {
    std::make_heap(ages.begin(), ages.end());
    while ((int64_t)ages.size() > maxCount) {
      release(ages.front().iQuiz);
      std::pop_heap(ages.begin(), ages.end());
      ages.pop_back();
    }
  }
  residentN_ = 0;
  return nullptr;
}

// BaseEngine::StartMaintenance (BaseEngine.cpp:640-683): with quizzes alive it either destroys them all
// (forceQuizzes) or refuses with QuizzesActive and stays in regular mode.
PqaError *Engine::StartMaintenance(bool forceQuizzes) {
  if (IsSharded()) return ErrNotImplemented("maintenance mode on a sharded engine");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (maintenance_) return MakeError(ErrCode::MaintenanceModeAlreadyThis, PQA_FILE_LINE "The engine is in maintenance mode already.",
                                     "activeMode=1");
  const int64_t nQuizzes = (int64_t)quizzes_.size() - (int64_t)quizGaps_.size();
  if (nQuizzes != 0) {
    if (!forceQuizzes)
      return MakeError(ErrCode::QuizzesActive, PQA_FILE_LINE "Can't switch to maintenance mode because there are active quizzes"
                       " and forceQuizzes=false", "nQuizzes=" + std::to_string(nQuizzes));
    for (int64_t i = 0; i < (int64_t)quizzes_.size(); i++) {
      HostQuiz &q = quizzes_[(size_t)i];
      if (!q.present) continue;
      q.present = false; q.answers.clear(); q.answers.shrink_to_fit(); q.activeQuestion = -1;
      quizGaps_.push_back(i);                       // :661
      pimQuiz_.RemoveComp(i);                       // :662
    }
  }
  residentN_ = 0;
  maintenance_ = true;
  return nullptr;
}

// BaseEngine::FinishMaintenance (BaseEngine.cpp:685-702)
PqaError *Engine::FinishMaintenance() {
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  if (!maintenance_) return MakeError(ErrCode::MaintenanceModeAlreadyThis, PQA_FILE_LINE "The engine is in regular mode already.",
                                      "activeMode=0");
  PQA_TRY
  SyncGapBits();
  maintenance_ = false;
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// New device arrays for newQ x K x newT, old cells kept; the cells of new rows / columns are written by the caller.
// The quiz pool (sized by Tp and ceil(Q/64)) is dropped: maintenance mode has no quizzes.
void Engine::ResizeKB(int64_t newQ, int64_t newT) {
  MarkKBChanged();
  const int64_t newTp = (newT + 3) & ~3ll;
  double *nA = nullptr, *nD = nullptr, *nB = nullptr;
  PQA_CU(cudaMalloc(&nA, sizeof(double) * (size_t)(newQ * K_ * newTp)));
  PQA_CU(cudaMalloc(&nD, sizeof(double) * (size_t)(newQ * newTp)));
  PQA_CU(cudaMalloc(&nB, sizeof(double) * (size_t)newTp));
  launch_copy_rows(nA, newTp, dSA_, Tp_, Q_ * K_, T_, stream_);
  launch_copy_rows(nD, newTp, dMD_, Tp_, Q_, T_, stream_);
  launch_copy_rows(nB, newTp, dVB_, Tp_, 1, T_, stream_);
  // padding lanes of every row: sA = 0, mD = 1, vB = 0 (pqa_kernels.cuh DeviceKB)
  if (newTp > newT) {
    std::vector<FillRect> pads = {FillRect{nA, newTp, 0, newQ * K_, newT, newTp - newT, 0.0},
                                  FillRect{nD, newTp, 0, newQ, newT, newTp - newT, 1.0},
                                  FillRect{nB, newTp, 0, 1, newT, newTp - newT, 0.0}};
    DevBuf<FillRect> dPads;
    dPads.ensure(pads.size(), stream_);
    PQA_CU(cudaMemcpyAsync(dPads.get(), pads.data(), sizeof(FillRect) * pads.size(), cudaMemcpyHostToDevice, stream_));
    launch_fill_rects(dPads.get(), (int64_t)pads.size(), stream_);
    PQA_CU(cudaStreamSynchronize(stream_));
  }
  PQA_CU(cudaStreamSynchronize(stream_));
  cudaFree(dSA_); cudaFree(dMD_); cudaFree(dVB_);
  dSA_ = nA; dMD_ = nD; dVB_ = nB;
  DropQuizPool();
  Q_ = newQ; T_ = newT; Tp_ = newTp;
  qLocal_ = Q_; tLocal_ = T_; TpL_ = Tp_;
  askedWords_ = (Q_ + 63) >> 6;
}

// BaseEngine::AddQsTs (BaseEngine.cpp:704-719) -> CpuEngine::AddQsTsSpec (CpuEngine.cpp:468-583). Gaps are reused first
// (LIFO), the rest is appended. Initial amounts: a question row gets init^2 in every sA cell and K*init^2 in mD; a
// target column gets its own init^2 / K*init^2 in the rows of the questions that existed before the call (rows of
// questions added or re-added by the same call keep the question's amounts), and vB = init.
// Note: the reference indexes the parameters of appended targets with the number of REUSED QUESTIONS
// (`pAtps[nQReuse + j]`, CpuEngine.cpp:517,523,528) where the number of reused targets is meant; the two agree whenever
// nQReuse == nTReuse (in particular with no gaps) and the reference reads the wrong / out-of-range entry otherwise.
// This engine uses the entry the call was given for that target.
PqaError *Engine::AddQsTs(int64_t nQuestions, CiAddQorTParam *pAqps, int64_t nTargets, CiAddQorTParam *pAtps) {
  if (!maintenance_) return NotMaintenance("add questions/targets");
  if (nQuestions < 0) return ErrNegativeCount(nQuestions, PQA_FILE_LINE "|nQuestions| must be non-negative.");
  if (nTargets < 0) return ErrNegativeCount(nTargets, PQA_FILE_LINE "|nTargets| must be non-negative.");
  if ((nQuestions > 0 && !pAqps) || (nTargets > 0 && !pAtps)) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pAqps/pAtps");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  const int64_t nQReuse = std::min(nQuestions, qGaps_.GetNGaps());
  const int64_t nQNew = nQuestions - nQReuse, nQOld = Q_, totQ = nQOld + nQNew;
  std::vector<uint8_t> reusedQ((size_t)totQ, 0);
  for (int64_t i = 0; i < nQReuse; i++) {                                  // CpuEngine.cpp:477-483
    const int64_t q = qGaps_.Acquire();
    pimQ_.RenewComp(q);
    pAqps[i]._index = q;
    reusedQ[(size_t)q] = 1;
  }
  const int64_t nTReuse = std::min(nTargets, tGaps_.GetNGaps());
  const int64_t nTNew = nTargets - nTReuse, nTOld = T_, totT = nTOld + nTNew;
  for (int64_t i = 0; i < nTReuse; i++) {                                  // :489-494
    const int64_t t = tGaps_.Acquire();
    pimT_.RenewComp(t);
    pAtps[i]._index = t;
  }
  for (int64_t i = 0; i < nQNew; i++) pAqps[nQReuse + i]._index = nQOld + i;   // :501
  for (int64_t j = 0; j < nTNew; j++) pAtps[nTReuse + j]._index = nTOld + j;   // :535
  if (nQNew > 0 || nTNew > 0) ResizeKB(totQ, totT);
  const int64_t Tp = Tp_;
  const double dK = (double)K_;
  // phase A: rows of appended questions, columns of appended and re-added targets in the rows of the old questions
  std::vector<FillRect> a, b;
  for (int64_t i = 0; i < nQNew; i++) {                                    // :498-512
    const double init = pAqps[nQReuse + i]._initAmount, sq = init * init;
    a.push_back(FillRect{dSA_, Tp, (nQOld + i) * K_, K_, 0, totT, sq});
    a.push_back(FillRect{dMD_, Tp, nQOld + i, 1, 0, totT, sq * dK});
  }
  for (int64_t j = 0; j < nTNew; j++) {                                    // :514-538
    const double init = pAtps[nTReuse + j]._initAmount, sq = init * init;
    a.push_back(FillRect{dSA_, Tp, 0, nQOld * K_, nTOld + j, 1, sq});
    a.push_back(FillRect{dMD_, Tp, 0, nQOld, nTOld + j, 1, sq * dK});
    a.push_back(FillRect{dVB_, Tp, 0, 1, nTOld + j, 1, init});
  }
  for (int64_t j = 0; j < nTReuse; j++) {                                  // :563-578
    const double init = pAtps[j]._initAmount, sq = init * init;
    a.push_back(FillRect{dSA_, Tp, 0, nQOld * K_, pAtps[j]._index, 1, sq});
    a.push_back(FillRect{dMD_, Tp, 0, nQOld, pAtps[j]._index, 1, sq * dK});
    a.push_back(FillRect{dVB_, Tp, 0, 1, pAtps[j]._index, 1, init});
  }
  // phase B: rows of re-added questions win over the target columns written in phase A (:552-562, and the skip at :568-570)
  for (int64_t i = 0; i < nQReuse; i++) {
    const double init = pAqps[i]._initAmount, sq = init * init;
    b.push_back(FillRect{dSA_, Tp, pAqps[i]._index * K_, K_, 0, totT, sq});
    b.push_back(FillRect{dMD_, Tp, pAqps[i]._index, 1, 0, totT, sq * dK});
  }
  DevBuf<FillRect> dRects;
  for (std::vector<FillRect> *phase : {&a, &b}) {
    if (phase->empty()) continue;
    dRects.ensure(phase->size(), stream_);
    PQA_CU(cudaMemcpyAsync(dRects.get(), phase->data(), sizeof(FillRect) * phase->size(), cudaMemcpyHostToDevice, stream_));
    launch_fill_rects(dRects.get(), (int64_t)phase->size(), stream_);
    PQA_CU(cudaStreamSynchronize(stream_));
  }
  pimQ_.GrowTo(totQ); pimT_.GrowTo(totT);                                  // :544-548
  qGaps_.GrowTo(totQ); tGaps_.GrowTo(totT);
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

// BaseEngine::RemoveQuestions / RemoveTargets (BaseEngine.cpp:721-766): the ids become gaps; cells stay in place.
PqaError *Engine::RemoveQuestions(int64_t nQuestions, const int64_t *pQIds) {
  if (!maintenance_) return NotMaintenance("remove questions");
  if (nQuestions > 0 && !pQIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pQIds");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t i = 0; i < nQuestions; i++) {
    const int64_t q = pQIds[i];
    if (q < 0 || q >= Q_ || qGaps_.IsGap(q)) return ErrAbsentId(q, PQA_FILE_LINE "Question index is not in KB.");
    qGaps_.Release(q);
    pimQ_.RemoveComp(q);
  }
  return nullptr;
}
PqaError *Engine::RemoveTargets(int64_t nTargets, const int64_t *pTIds) {
  if (!maintenance_) return NotMaintenance("remove targets");
  if (nTargets > 0 && !pTIds) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "pTIds");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  for (int64_t i = 0; i < nTargets; i++) {
    const int64_t t = pTIds[i];
    if (t < 0 || t >= T_ || tGaps_.IsGap(t)) return ErrAbsentId(t, PQA_FILE_LINE "Target index is not in KB (but rather at a gap).");
    tGaps_.Release(t);
    pimT_.RemoveComp(t);
  }
  return nullptr;
}

// The old-id arrays of CpuEngine::CompactSpec. Questions (CpuEngine.cpp:594-608): scanning from the front, every gap takes
// the last non-gap question. Targets (:616-641): the same scan records the gap positions (ascending) and the moved targets
// (descending) and pairs the g-th smallest gap with the g-th smallest moved target.
void PlanCompactQuestions(const GapSet &gaps, int64_t n, int64_t *oldIds) {
  int64_t iFirst, iLast;
  for (iFirst = 0, iLast = n - 1; iFirst <= iLast; iFirst++) {
    if (!gaps.IsGap(iFirst)) { oldIds[iFirst] = iFirst; continue; }
    while (gaps.IsGap(iLast) && iLast > iFirst) iLast--;
    if (iFirst == iLast) break;
    oldIds[iFirst] = iLast;
    iLast--;
  }
}
void PlanCompactTargets(const GapSet &gaps, int64_t n, int64_t *oldIds) {
  std::vector<int64_t> dests, srcs;
  int64_t iFirst, iLast;
  for (iFirst = 0, iLast = n - 1; iFirst <= iLast; iFirst++) {
    if (!gaps.IsGap(iFirst)) { oldIds[iFirst] = iFirst; continue; }
    while (gaps.IsGap(iLast) && iLast > iFirst) iLast--;
    if (iFirst == iLast) break;
    dests.push_back(iFirst);
    srcs.push_back(iLast);
    iLast--;
  }
  const size_t nMoves = dests.size();
  for (size_t g = 0; g < nMoves; g++) oldIds[dests[g]] = srcs[nMoves - 1 - g];
}

// Host-logic self test (no device needed; called by the CPU test suite through PqaB200_HostLogicSelfTest): gap sets,
// permanent ids and the compaction planners against brute-force expectations. Returns an empty string when all is well.
std::string HostLogicSelfTest() {
  auto fail = [](const char *what, int64_t a, int64_t b) { return std::string(what) + " (" + std::to_string(a) + " vs " + std::to_string(b) + ")"; };
  uint64_t lcg = 88172645463325252ull;
  auto rnd = [&](uint64_t m) { lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17; return (int64_t)(lcg % m); };
  for (int trial = 0; trial < 200; trial++) {
    const int64_t n = 2 + rnd(60);
    GapSet gs; gs.GrowTo(n);
    PermIds pim; pim.GrowTo(n);
    std::vector<int64_t> removedOrder;
    std::vector<uint8_t> gone((size_t)n, 0);
    const int64_t nRemove = rnd((uint64_t)n);                 // at least one id survives
    for (int64_t r = 0; r < nRemove; r++) {
      const int64_t id = rnd((uint64_t)n);
      if (gone[(size_t)id]) continue;
      gone[(size_t)id] = 1; removedOrder.push_back(id);
      gs.Release(id);
      if (!pim.RemoveComp(id)) return "RemoveComp refused a live id";
      if (pim.RemoveComp(id)) return "RemoveComp accepted a removed id";
    }
    if (gs.GetNGaps() != (int64_t)removedOrder.size()) return fail("gap count", gs.GetNGaps(), (int64_t)removedOrder.size());
    for (int64_t i = 0; i < n; i++) {
      if (gs.IsGap(i) != (gone[(size_t)i] != 0)) return fail("IsGap", i, gone[(size_t)i]);
      if (pim.PermFromComp(i) != (gone[(size_t)i] ? -1 : i)) return fail("PermFromComp", i, pim.PermFromComp(i));
      if (pim.CompFromPerm(i) != (gone[(size_t)i] ? -1 : i)) return fail("CompFromPerm", i, pim.CompFromPerm(i));
    }
    // compaction plans: a permutation of the survivors onto 0..nLive-1, fixed points stay, questions take the LAST
    // survivor for the first gap, targets pair ascending gaps with ascending moved survivors
    const int64_t nLive = n - gs.GetNGaps();
    std::vector<int64_t> oq((size_t)n, -7), ot((size_t)n, -7);
    PlanCompactQuestions(gs, n, oq.data());
    PlanCompactTargets(gs, n, ot.data());
    for (const std::vector<int64_t> *plan : {&oq, &ot}) {
      std::vector<uint8_t> used((size_t)n, 0);
      for (int64_t i = 0; i < nLive; i++) {
        const int64_t o = (*plan)[(size_t)i];
        if (o < 0 || o >= n || gone[(size_t)o] || used[(size_t)o]) return fail("compaction plan is not a permutation of the survivors", i, o);
        used[(size_t)o] = 1;
        if (!gone[(size_t)i] && o != i) return fail("a surviving id below the new size moved", i, o);
        if (gone[(size_t)i] && o < nLive) return fail("a gap was filled from inside the new range", i, o);
      }
    }
    int64_t prevQ = n, prevT = -1;
    for (int64_t i = 0; i < nLive; i++) {
      if (!gone[(size_t)i]) continue;
      if (oq[(size_t)i] >= prevQ) return fail("questions: gaps must take survivors from the end downwards", i, oq[(size_t)i]);
      if (ot[(size_t)i] <= prevT) return fail("targets: gaps must take moved survivors in ascending order", i, ot[(size_t)i]);
      prevQ = oq[(size_t)i]; prevT = ot[(size_t)i];
    }
    // permanent ids follow their rows through the compaction; re-added ids get fresh permanent ids
    PermIds pim2 = pim;
    if (!pim2.OnCompact(nLive, oq.data())) return "OnCompact refused a valid plan";
    for (int64_t i = 0; i < nLive; i++)
      if (pim2.PermFromComp(i) != oq[(size_t)i] || pim2.CompFromPerm(oq[(size_t)i]) != i) return fail("OnCompact mapping", i, pim2.PermFromComp(i));
    int64_t nextPerm = n;
    while (gs.GetNGaps() > 0) {
      const int64_t want = removedOrder.back(); removedOrder.pop_back();
      const int64_t got = gs.Acquire();                        // LIFO reuse (GapTracker.h:38-49)
      if (got != want) return fail("gap reuse order", got, want);
      if (!pim.RenewComp(got) || pim.PermFromComp(got) != nextPerm) return fail("RenewComp permanent id", pim.PermFromComp(got), nextPerm);
      nextPerm++;
    }
    if (gs.Acquire() != n) return "Acquire without gaps must append";
  }
  PermIds q;
  q.GrowTo(3);
  if (q.EnsurePermIdGreater(1) || !q.EnsurePermIdGreater(9)) return "EnsurePermIdGreater";
  q.GrowTo(4);
  if (q.PermFromComp(3) != 10) return "GrowTo after EnsurePermIdGreater";
  if (q.RemapPermId(0, 10) || q.RemapPermId(0, 11) || !q.RemapPermId(0, 5) || q.CompFromPerm(5) != 0 || q.CompFromPerm(0) != -1) return "RemapPermId";
  return "";
}

// BaseEngine::Compact (BaseEngine.cpp:768-779) -> CpuEngine::CompactSpec (CpuEngine.cpp:585-658).
// Questions: scanning from the front, every gap takes the last non-gap question (rows swapped, :594-608).
// Targets: the same scan records the gap positions in ascending order and the moved targets in descending order, and
// then pairs the g-th smallest gap position with the g-th smallest moved target (moves[] is filled from both ends,
// :620-633). When removed targets sit among the last nGaps positions the reference's loop ends before moves[] is full
// and its copy loop reads unset entries (it asserts in debug builds); here only the recorded moves are applied, paired
// the same way.
PqaError *Engine::Compact(int64_t *pnQuestions, const int64_t **ppOldQuestions, int64_t *pnTargets,
                          const int64_t **ppOldTargets) {
  if (pnQuestions) *pnQuestions = 0;
  if (pnTargets) *pnTargets = 0;
  if (ppOldQuestions) *ppOldQuestions = nullptr;
  if (ppOldTargets) *ppOldTargets = nullptr;
  if (!maintenance_) return NotMaintenance("compact the KB");
  if (!pnQuestions || !ppOldQuestions || !pnTargets || !ppOldTargets) return MakeError(ErrCode::NullArgument, PQA_FILE_LINE "out-pointers");
  std::lock_guard<std::mutex> lk(mu_); DeviceScope devScope(device_);
  PQA_TRY
  const int64_t nQ = Q_ - qGaps_.GetNGaps(), nT = T_ - tGaps_.GetNGaps();
  if (nQ < 1 || nT < 2) return ErrInsufficientDims(K_, nQ, nT);
  std::unique_ptr<int64_t[]> oldQ(new int64_t[(size_t)std::max<int64_t>(nQ, 1)]), oldT(new int64_t[(size_t)std::max<int64_t>(nT, 1)]);
  PlanCompactQuestions(qGaps_, Q_, oldQ.get());
  PlanCompactTargets(tGaps_, T_, oldT.get());
  const int64_t newTp = (nT + 3) & ~3ll;
  DevBuf<int64_t> dOldQ, dOldT, dZero;
  dOldQ.ensure((size_t)nQ, stream_); dOldT.ensure((size_t)nT, stream_); dZero.ensure(1, stream_);
  const int64_t zero = 0;
  PQA_CU(cudaMemcpyAsync(dOldQ.get(), oldQ.get(), sizeof(int64_t) * (size_t)nQ, cudaMemcpyHostToDevice, stream_));
  PQA_CU(cudaMemcpyAsync(dOldT.get(), oldT.get(), sizeof(int64_t) * (size_t)nT, cudaMemcpyHostToDevice, stream_));
  PQA_CU(cudaMemcpyAsync(dZero.get(), &zero, sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
  double *nA = nullptr, *nD = nullptr, *nB = nullptr;
  PQA_CU(cudaMalloc(&nA, sizeof(double) * (size_t)(nQ * K_ * newTp)));
  PQA_CU(cudaMalloc(&nD, sizeof(double) * (size_t)(nQ * newTp)));
  PQA_CU(cudaMalloc(&nB, sizeof(double) * (size_t)newTp));
  launch_gather_kb(nA, newTp, dSA_, Tp_, dOldQ.get(), nQ, K_, dOldT.get(), nT, 0.0, stream_);
  launch_gather_kb(nD, newTp, dMD_, Tp_, dOldQ.get(), nQ, 1, dOldT.get(), nT, 1.0, stream_);
  launch_gather_kb(nB, newTp, dVB_, Tp_, dZero.get(), 1, 1, dOldT.get(), nT, 0.0, stream_);
  PQA_CU(cudaStreamSynchronize(stream_));
  cudaFree(dSA_); cudaFree(dMD_); cudaFree(dVB_);
  dSA_ = nA; dMD_ = nD; dVB_ = nB;
  DropQuizPool();
  qGaps_.Compact(nQ); tGaps_.Compact(nT);                                  // :609, :651
  pimQ_.OnCompact(nQ, oldQ.get()); pimT_.OnCompact(nT, oldT.get());        // :611, :653
  Q_ = nQ; T_ = nT; Tp_ = newTp; qLocal_ = Q_; tLocal_ = T_; TpL_ = Tp_;
  askedWords_ = (Q_ + 63) >> 6;
  *pnQuestions = nQ; *pnTargets = nT;
  *ppOldQuestions = oldQ.release(); *ppOldTargets = oldT.release();        // released by CiReleaseCompaction (delete[])
  return nullptr;
  PQA_CATCH_RETURN_ERR
}

} // namespace pqa
