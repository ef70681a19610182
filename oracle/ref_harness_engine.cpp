// TEST INFRASTRUCTURE: C-ABI harness that drives the REFERENCE's own hot-path subtask bodies
// (CEEvalQsSubtaskConsider::Run, CERecordAnswerSubtaskMul::Run, CESetPriorsSubtaskSum::Run,
// CEDivTargPriorsSubtask::Run, CEHeapifyPriorsSubtaskMake::Run + CEListTopTargetsAlgorithm::RunHeapifyBased,
// CETrainOperation::Perform1/2) compiled unmodified-in-logic from /root/reference/ProbQA (scratch copies,
// MSVC->GCC token patches only, see build_ref.sh) against the stub engine shell in refshim/stubs/.
// The few driver lines restated here cite the reference lines they mirror. Output: oracle/_ref/libpqa_ref.so.
#include "stdafx.h"
#include "../PqaCore/CpuEngine.h"
#include "../PqaCore/CEQuiz.h"
#include "../PqaCore/CEEvalQsTask.h"
#include "../PqaCore/CEEvalQsSubtaskConsider.h"
#include "../PqaCore/CERecordAnswerTask.h"
#include "../PqaCore/CERecordAnswerSubtaskMul.h"
#include "../PqaCore/CESetPriorsTask.h"
#include "../PqaCore/CESetPriorsSubtaskSum.h"
#include "../PqaCore/CEDivTargPriorsSubtask.h"
#include "../PqaCore/Summator.h"
#include "../PqaCore/CEListTopTargetsAlgorithm.h"
#include "../PqaCore/CETrainOperation.h"
#include "../PqaCore/CETrainTaskNumSpec.h"
#include "../PqaCore/CEUpdatePriorsTask.h"
#include "../PqaCore/CEUpdatePriorsSubtaskMul.h"
#include "../PqaCore/CENormPriorsTask.h"
#include "../PqaCore/CENormPriorsSubtaskMax.h"
#include "../PqaCore/CENormPriorsSubtaskCorrSum.h"

using namespace SRPlat;
using namespace ProbQA;

// Out-of-line members of PqaError that the reference defines in PqaErrors.cpp (string plumbing, not compiled here).
namespace ProbQA {
PqaError::~PqaError() {}
}

namespace {
typedef CpuEngine<SRDoubleNumber> TEngine;
typedef CEQuiz<SRDoubleNumber> TQuiz;

struct RefEngine {
  TEngine eng;
  TQuiz quiz;
  RefEngine(const EngineDimensions& d, SRThreadCount w) : eng(d, w), quiz(d) {}
};

void load_prior(RefEngine *re, const double *prior) {
  const size_t T = size_t(re->eng._dims._nTargets);
  double *p = reinterpret_cast<double*>(re->quiz.GetPriorMants());
  for (size_t j = 0; j < re->quiz._ld; j++) p[j] = (j < T) ? prior[j] : 0.0;
}
void store_prior(RefEngine *re, double *prior) {
  memcpy(prior, re->quiz.GetPriorMants(), sizeof(double) * size_t(re->eng._dims._nTargets));
}
void load_asked(RefEngine *re, const uint8_t *asked) {
  memset(re->quiz.GetQAsked(), 0, 32 * re->quiz._nAskedVects);
  if (asked) memcpy(re->quiz.GetQAsked(), asked, size_t((re->eng._dims._nQuestions + 7) / 8));
}
} // anonymous namespace

extern "C" {

// sA/mD/vB are flat with row stride T (the reference's file layout). nWorkers plays hardware_concurrency().
void* ref_engine_create(int64_t Q, int64_t K, int64_t T, const double *sA, const double *mD, const double *vB,
                        const uint8_t *qgaps, const uint8_t *tgaps, int64_t nWorkers) {
  EngineDimensions d; d._nAnswers = K; d._nQuestions = Q; d._nTargets = T;
  RefEngine *re = new RefEngine(d, SRThreadCount(nWorkers));
  TEngine &e = re->eng;
  double *pA = reinterpret_cast<double*>(e._sA), *pD = reinterpret_cast<double*>(e._mD), *pB = reinterpret_cast<double*>(e._vB);
  for (int64_t r = 0; r < Q * K; r++)
    for (size_t j = 0; j < e._ld; j++) pA[size_t(r) * e._ld + j] = (int64_t(j) < T) ? sA[r * T + int64_t(j)] : 0.0;
  for (int64_t r = 0; r < Q; r++)
    for (size_t j = 0; j < e._ld; j++) pD[size_t(r) * e._ld + j] = (int64_t(j) < T) ? mD[r * T + int64_t(j)] : 1.0;
  for (size_t j = 0; j < e._ld; j++) pB[j] = (int64_t(j) < T) ? vB[j] : 0.0;
  e._questionGaps.Assign(qgaps, Q);
  e._targetGaps.Assign(tgaps, T);
  return re;
}
void ref_engine_destroy(void *h) { delete static_cast<RefEngine*>(h); }
void ref_engine_set_os_threads(void *h, int64_t n) { static_cast<RefEngine*>(h)->eng._tpWorkers.SetOsThreads(SRThreadCount(n)); }
long long ref_log_count() { return RefShimLogCount().load(); }

void ref_engine_read_kb(void *h, double *sA, double *mD, double *vB) {
  RefEngine *re = static_cast<RefEngine*>(h); TEngine &e = re->eng;
  const int64_t Q = e._dims._nQuestions, K = e._dims._nAnswers, T = e._dims._nTargets;
  const double *pA = reinterpret_cast<const double*>(e._sA), *pD = reinterpret_cast<const double*>(e._mD), *pB = reinterpret_cast<const double*>(e._vB);
  for (int64_t r = 0; r < Q * K; r++) memcpy(sA + r * T, pA + size_t(r) * e._ld, sizeof(double) * size_t(T));
  for (int64_t r = 0; r < Q; r++) memcpy(mD + r * T, pD + size_t(r) * e._ld, sizeof(double) * size_t(T));
  memcpy(vB, pB, sizeof(double) * size_t(T));
}

// CECreateQuizStart::UpdateLikelihoods, CECreateQuizOperation.cpp:22-53
void ref_start_quiz(void *h, double *priorOut) {
  RefEngine *re = static_cast<RefEngine*>(h); TEngine &engine = re->eng; TQuiz &quiz = re->quiz;
  const EngineDimensions& dims = engine.GetDims();
  const SRThreadCount nWorkers = engine.GetWorkers().GetWorkerCount();              // :29
  std::vector<uint8_t> stMem(nWorkers * SRMaxSizeof<CESetPriorsSubtaskSum<SRDoubleNumber>,
    CEDivTargPriorsSubtask<CESetPriorsTask<SRDoubleNumber>>>::value + 64);
  std::vector<size_t> splitMem(nWorkers + 1);
  SRPoolRunner pr(engine.GetWorkers(), stMem.data());
  const TPqaId nTargetVects = SRSimd::VectsFromComps<SRDoubleNumber>(dims._nTargets); // :39
  const SRPoolRunner::Split targSplit = SRPoolRunner::CalcSplit(splitMem.data(), nTargetVects, nWorkers); // :40
  CESetPriorsTask<SRDoubleNumber> spTask(engine, quiz);
  {
    typedef CESetPriorsSubtaskSum<SRDoubleNumber> TSubtask;
    SRPoolRunner::Keeper<TSubtask> kp = pr.RunPreSplit<TSubtask>(spTask, targSplit);  // :47
    Summator<SRDoubleNumber>::ForPriors(kp, spTask);                                  // :49
  }
  pr.RunPreSplit<CEDivTargPriorsSubtask<CESetPriorsTask<SRDoubleNumber>>>(spTask, targSplit); // :52
  store_prior(re, priorOut);
}

// CEQuiz::RecordAnswer, CEQuiz.h:94-121 (the answers vector / asked bit / validation are host bookkeeping)
void ref_record_answer(void *h, double *prior, int64_t iQuestion, int64_t iAnswer) {
  RefEngine *re = static_cast<RefEngine*>(h); TEngine &engine = re->eng; TQuiz &quiz = re->quiz;
  load_prior(re, prior);
  const EngineDimensions& dims = engine.GetDims();
  const SRThreadCount nWorkers = engine.GetNLooseWorkers();                           // :98
  std::vector<uint8_t> stMem(nWorkers * SRMaxSizeof<CERecordAnswerSubtaskMul<SRDoubleNumber>,
    CEDivTargPriorsSubtask<CERecordAnswerTask<SRDoubleNumber>>>::value + 64);
  std::vector<size_t> splitMem(nWorkers + 1);
  SRPoolRunner pr(engine.GetWorkers(), stMem.data());
  const TPqaId nTargetVects = SRSimd::VectsFromComps<SRDoubleNumber>(dims._nTargets); // :109
  const SRPoolRunner::Split targSplit = SRPoolRunner::CalcSplit(splitMem.data(), nTargetVects, nWorkers); // :110
  const AnsweredQuestion aq(iQuestion, iAnswer);
  CERecordAnswerTask<SRDoubleNumber> raTask(engine, quiz, aq);                        // :112
  {
    typedef CERecordAnswerSubtaskMul<SRDoubleNumber> TSubtask;
    SRPoolRunner::Keeper<TSubtask> kp = pr.RunPreSplit<TSubtask>(raTask, targSplit);  // :116
    Summator<SRDoubleNumber>::ForPriors(kp, raTask);                                  // :117
  }
  pr.RunPreSplit<CEDivTargPriorsSubtask<CERecordAnswerTask<SRDoubleNumber>>>(raTask, targSplit); // :120
  store_prior(re, prior);
}

// CpuEngine::NextQuestionSpec, CpuEngine.cpp:337-374: runs the evaluation subtasks over split(Q, 8*hwc) and
// builds the grand totals. Returns the number of chunks. runLength[Q], grand[8W], bounds[8W].
int64_t ref_eval_questions(void *h, const double *prior, const uint8_t *asked, double *runLength, double *grand,
                           int64_t *bounds) {
  RefEngine *re = static_cast<RefEngine*>(h); TEngine &engine = re->eng; TQuiz &quiz = re->quiz;
  load_prior(re, prior); load_asked(re, asked);
  const EngineDimensions& dims = engine.GetDims();
  const SRSubtaskCount nWorkers = engine.GetWorkers().GetWorkerCount() * 8;           // :339
  std::vector<uint8_t> stMem(nWorkers * SRMaxSizeof<CEEvalQsSubtaskConsider<SRDoubleNumber>>::value + 64);
  std::vector<size_t> splitMem(nWorkers + 1);
  SRPoolRunner pr(engine.GetWorkers(), stMem.data());
  SRDoubleNumber *pRun = reinterpret_cast<SRDoubleNumber*>(runLength);
  CEEvalQsTask<SRDoubleNumber> evalQsTask(engine, quiz, dims._nTargets - engine.GetTargetGaps().GetNGaps(), pRun); // :351
  const SRPoolRunner::Split questionSplit = SRPoolRunner::CalcSplit(splitMem.data(), dims._nQuestions, nWorkers); // :354
  {
    SRPoolRunner::Keeper<CEEvalQsSubtaskConsider<SRDoubleNumber>> kp =
      pr.RunPreSplit<CEEvalQsSubtaskConsider<SRDoubleNumber>>(evalQsTask, questionSplit); // :359
  }
  SRAccumulator<SRDoubleNumber> accTotG(SRDoubleNumber(0.0));                         // :362
  for (SRSubtaskCount i = 0; i < questionSplit._nSubtasks; i++) {                     // :365-369
    const SRDoubleNumber curGT = pRun[questionSplit._pBounds[i] - 1];
    accTotG.Add(curGT);
    grand[i] = accTotG.Get().GetValue();
    bounds[i] = int64_t(questionSplit._pBounds[i]);
  }
  return int64_t(questionSplit._nSubtasks);
}

// ResumeQuiz: CECreateQuizResume::UpdateLikelihoods (CECreateQuizOperation.cpp:55-83) driving the reference's
// CEUpdatePriorsSubtaskMul::Run, then CpuEngine::NormalizePriors (CpuEngine.cpp:284-335) driving
// CENormPriorsSubtaskMax::Run, CENormPriorsSubtaskCorrSum::Run and CEDivTargPriorsSubtask::Run.
// Returns 0, or 1 for the reference's I64Underflow error (:315-318).
int64_t ref_resume_quiz(void *h, const AnsweredQuestion *pAQs, int64_t nAnswered, double *priorOut) {
  RefEngine *re = static_cast<RefEngine*>(h); TEngine &engine = re->eng; TQuiz &quiz = re->quiz;
  const EngineDimensions& dims = engine.GetDims();
  const SRThreadCount nWorkers = engine.GetWorkers().GetWorkerCount();                // CECreateQuizOperation.cpp:63
  std::vector<uint8_t> stMem(nWorkers * 256 + 64);
  std::vector<size_t> splitMem(nWorkers + 1);
  SRPoolRunner pr(engine.GetWorkers(), stMem.data());
  const TPqaId nTargetVects = SRSimd::VectsFromComps<SRDoubleNumber>(dims._nTargets); // :73
  const SRPoolRunner::Split targSplit = SRPoolRunner::CalcSplit(splitMem.data(), nTargetVects, nWorkers); // :74
  {
    CEUpdatePriorsTask<SRDoubleNumber> task(engine, quiz, nAnswered, pAQs, 0 /* no cache blocking: same arithmetic */);
    pr.RunPreSplit<CEUpdatePriorsSubtaskMul<SRDoubleNumber>>(task, targSplit);        // :79
  }
  CENormPriorsTask<SRDoubleNumber> normPriorsTask(engine, quiz);                      // CpuEngine.cpp:287
  {
    SRPoolRunner::Keeper<CENormPriorsSubtaskMax<SRDoubleNumber>> kp =
      pr.RunPreSplit<CENormPriorsSubtaskMax<SRDoubleNumber>>(normPriorsTask, targSplit); // :290-291
    int64_t fullMax = std::numeric_limits<int64_t>::min();                            // :293-313 is an integer max
    for (SRSubtaskCount i = 0; i < kp.GetNSubtasks(); i++) fullMax = std::max(fullMax, kp.GetSubtask(i)->_maxExp);
    const int64_t highBound = SRDoubleNumber::_cMaxExp + SRDoubleNumber::_cExpOffs - SRMath::CeilLog2(dims._nTargets) - 2; // :314
    const int64_t minAllowed = std::numeric_limits<int64_t>::min() + highBound + 1;
    if (fullMax <= minAllowed) return 1;                                              // :316-319
    normPriorsTask._corrExp = _mm256_set1_epi64x(highBound - fullMax);                // :320
  }
  {
    typedef CENormPriorsSubtaskCorrSum<SRDoubleNumber> TCorrSumSubtask;
    SRPoolRunner::Keeper<TCorrSumSubtask> kp = pr.RunPreSplit<TCorrSumSubtask>(normPriorsTask, targSplit); // :325
    Summator<SRDoubleNumber>::ForPriors(kp, normPriorsTask);                          // :326
  }
  pr.RunPreSplit<CEDivTargPriorsSubtask<CENormPriorsTask<SRDoubleNumber>>>(normPriorsTask, targSplit); // :330
  store_prior(re, priorOut);
  return 0;
}

// CpuEngine::ListTopTargetsSpec heapify branch, CpuEngine.cpp:417-440 -> CEListTopTargetsAlgorithm.cpp:30-97
int64_t ref_list_top_targets(void *h, const double *prior, int64_t maxCount, RatedTarget *dest) {
  RefEngine *re = static_cast<RefEngine*>(h);
  load_prior(re, prior);
  PqaError err;
  CEListTopTargetsAlgorithm<SRDoubleNumber> ltta(err, re->eng, re->quiz, maxCount, dest);
  return ltta.RunHeapifyBased();
}

// CpuEngine::RecordQuizTargetSpec, CpuEngine.cpp:442-466
void ref_record_quiz_target(void *h, const AnsweredQuestion *answers, int64_t nAnswers, int64_t iTarget, double amount) {
  RefEngine *re = static_cast<RefEngine*>(h);
  const CETrainTaskNumSpec<SRDoubleNumber> numSpec(amount);
  CETrainOperation<SRDoubleNumber> trainOp(re->eng, iTarget, numSpec);
  TPqaId i = 0;
  const TPqaId iEn = nAnswers - 1;
  for (; i < iEn; i += 2) trainOp.Perform2(answers[i], answers[i + 1]);               // :453-457
  if (i == iEn) trainOp.Perform1(answers[i]);                                         // :459-461
  re->eng.ModB(iTarget) += amount;                                                    // :462
}

} // extern "C"
