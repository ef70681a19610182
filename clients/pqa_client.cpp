// probqa_b200: the reference's learner driver (ProbQA/PqaClient/PqaClient.cpp:69-245, LearnerThread / LearnBinarySearch)
// re-expressed over the C ABI of libPqaCore.so -- only symbols of include/PqaCInterop.h are used, so this file also
// builds against the reference's own PqaCore. Learner threads run quizzes for random hidden targets on a
// 1000 x 5 x 1000 KB: StartQuiz, then up to 25 rounds of NextQuestion / RecordAnswer / ListTopTargets(1) until the hidden
// target is top-rated, then RecordQuizTarget and ReleaseQuiz. Every `--report-every` trainings one line goes to
// progress.txt in the reference's format (PqaClient.cpp:89-90):
//   nTrainings  totalQuestionsAsked  precision%  meanQuizLength  meanCertainty%  questionsAskedPerSecond
// The last column is the only throughput figure the reference publishes (mean 301 questions/s).
// The B200 engine combines the learners' concurrent one-quiz calls into batch launches, so more learner threads than
// cores are useful here (default 64).
#include <atomic>
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../include/PqaCInterop.h"

namespace {

constexpr int64_t kMaxQuizLen = 25;   // PqaClient.cpp:53
constexpr int64_t kTopRated = 1;      // :54

struct Shared {
  void *engine = nullptr;
  int64_t nTrainingsMax = 20000, reportEvery = 1024, nTargets = 1000, band = 32;
  std::string kbDir;
  std::mutex mu;                       // gcsReport
  int64_t nTrainings = -1, nCorrect = 0, nWrong = 0, sumQuizLens = 0;
  double totCertainty = 0;
  std::chrono::steady_clock::time_point start;
  uint64_t prevQAsked = 0;
  FILE *progress = nullptr;
  std::atomic<bool> failed{false};
  double lastRate = 0, lastPrecision = 0, sumRate = 0;
  int64_t nReports = 0;
};

std::string ErrText(void *err) {
  if (!err) return "";
  void *s = PqaError_ToString(err, 1);
  std::string out = s ? static_cast<const char *>(s) : "(no text)";
  CiReleaseString(s);
  CiReleasePqaError(err);
  return out;
}

bool Fail(Shared &sh, const char *what, void *err) {
  fprintf(stderr, "%s: %s\n", what, ErrText(err).c_str());
  sh.failed = true;
  return false;
}

void LearnerThread(Shared *psh, uint64_t seed) {
  Shared &sh = *psh;
  std::mt19937_64 rng(seed);
  while (!sh.failed) {
    {
      std::lock_guard<std::mutex> lk(sh.mu);
      sh.nTrainings++;
      if (sh.nTrainings > sh.nTrainingsMax) break;
      if (sh.nTrainings != 0 && sh.nTrainings % sh.reportEvery == 0) {      // PqaClient.cpp:78-106
        void *err = nullptr;
        const uint64_t totQAsked = PqaEngine_GetTotalQuestionsAsked(sh.engine, &err);
        if (err) { Fail(sh, "GetTotalQuestionsAsked", err); return; }
        const double precision = sh.nCorrect * 100.0 / (double)(sh.nCorrect + sh.nWrong);
        const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - sh.start).count();
        const double rate = (double)(totQAsked - sh.prevQAsked) / elapsed;
        fprintf(sh.progress, "%" PRId64 "\t%" PRIu64 "\t%lf\t%lf\t%lf\t%lf\n", sh.nTrainings, totQAsked, precision,
                (double)sh.sumQuizLens / (double)sh.nCorrect, sh.totCertainty / (double)sh.nCorrect, rate);
        fflush(sh.progress);
        if (!sh.kbDir.empty()) {
          char kbFile[512];
          snprintf(kbFile, sizeof(kbFile), "%s/dichotomy%.6" PRId64 ".kb", sh.kbDir.c_str(), sh.nTrainings);
          if (void *e = PqaEngine_SaveKB(sh.engine, kbFile, 0)) { Fail(sh, "SaveKB", e); return; }
        }
        sh.lastRate = rate; sh.lastPrecision = precision; sh.sumRate += rate; sh.nReports++;
        sh.nCorrect = sh.nWrong = sh.sumQuizLens = 0;
        sh.totCertainty = 0;
        sh.prevQAsked = totQAsked;
        sh.start = std::chrono::steady_clock::now();
      }
    }
    const int64_t guess = (int64_t)(rng() % (uint64_t)sh.nTargets);
    void *err = nullptr;
    const int64_t quiz = PqaEngine_StartQuiz(sh.engine, &err);
    if (err || quiz < 0) { Fail(sh, "StartQuiz", err); return; }
    int64_t j = 0;
    for (; j < kMaxQuizLen; j++) {
      const int64_t q = PqaEngine_NextQuestion(sh.engine, &err, quiz);
      if (err || q < 0) { Fail(sh, "NextQuestion", err); return; }
      int64_t a;                                                            // :118-137
      if (guess < q - sh.band) a = 0;
      else if (guess < q) a = 1;
      else if (guess == q) a = 2;
      else if (guess <= q + sh.band) a = 3;
      else a = 4;
      if (void *e = PqaEngine_RecordAnswer(sh.engine, quiz, a)) { Fail(sh, "RecordAnswer", e); return; }
      CiRatedTarget top[kTopRated];
      const int64_t nListed = PqaEngine_ListTopTargets(sh.engine, &err, quiz, kTopRated, top);
      if (err || nListed != kTopRated) { Fail(sh, "ListTopTargets", err); return; }
      if (top[0]._iTarget == guess) {
        std::lock_guard<std::mutex> lk(sh.mu);
        sh.nCorrect++;
        sh.sumQuizLens += j + 1;
        sh.totCertainty += top[0]._prob * 100;
        break;
      }
    }
    if (j >= kMaxQuizLen) { std::lock_guard<std::mutex> lk(sh.mu); sh.nWrong++; }
    if (void *e = PqaEngine_RecordQuizTarget(sh.engine, quiz, guess, 1.0)) { Fail(sh, "RecordQuizTarget", e); return; }
    if (void *e = PqaEngine_ReleaseQuiz(sh.engine, quiz)) { Fail(sh, "ReleaseQuiz", e); return; }
  }
}

}  // namespace

int main(int argc, char **argv) {
  Shared sh;
  int nLearners = 64;
  int64_t nQuestions = 1000;
  std::string initKb, progressPath = "progress.txt";
  for (int i = 1; i < argc; i++) {
    auto next = [&](const char *name) -> const char * {
      if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", name); exit(2); }
      return argv[++i];
    };
    if (!strcmp(argv[i], "--trainings")) sh.nTrainingsMax = atoll(next("--trainings"));
    else if (!strcmp(argv[i], "--learners")) nLearners = atoi(next("--learners"));
    else if (!strcmp(argv[i], "--report-every")) sh.reportEvery = atoll(next("--report-every"));
    else if (!strcmp(argv[i], "--size")) { nQuestions = atoll(next("--size")); sh.nTargets = nQuestions; }
    else if (!strcmp(argv[i], "--kb-dir")) sh.kbDir = next("--kb-dir");
    else if (!strcmp(argv[i], "--init-kb")) initKb = next("--init-kb");
    else if (!strcmp(argv[i], "--progress")) progressPath = next("--progress");
    else { fprintf(stderr, "usage: pqa_client [--trainings N] [--learners L] [--report-every R] [--size QT] [--kb-dir DIR] "
                           "[--init-kb FILE] [--progress FILE]\n"); return 2; }
  }
  sh.progress = fopen(progressPath.c_str(), "wt");
  if (!sh.progress) { perror("progress file"); return 1; }
  void *factory = CiGetPqaEngineFactory(), *err = nullptr;
  if (initKb.empty()) {                                                     // PqaClient.cpp:202-215
    CiEngineDefinition ed;
    memset(&ed, 0, sizeof(ed));
    ed._nAnswers = 5; ed._nQuestions = nQuestions; ed._nTargets = sh.nTargets;
    ed._initAmount = 0.1; ed._precType = 3 /* Double */; ed._memPoolMaxBytes = 512ull << 20;
    sh.engine = PqaEngineFactory_CreateCpuEngine(factory, &err, &ed);
  } else {
    sh.engine = PqaEngineFactory_LoadCpuEngine(factory, &err, initKb.c_str(), 512ull << 20);
  }
  if (err || !sh.engine) { fprintf(stderr, "Failed to instantiate a ProbQA engine: %s\n", ErrText(err).c_str()); return 1; }
  CiEngineDimensions dims;
  PqaEngine_CopyDims(sh.engine, &dims);
  sh.nTargets = dims._nTargets;
  sh.prevQAsked = PqaEngine_GetTotalQuestionsAsked(sh.engine, &err);
  sh.start = std::chrono::steady_clock::now();
  const auto t0 = sh.start;
  std::vector<std::thread> learners;
  for (int i = 0; i < nLearners; i++) learners.emplace_back(LearnerThread, &sh, 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
  for (auto &t : learners) t.join();
  const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const uint64_t asked = PqaEngine_GetTotalQuestionsAsked(sh.engine, &err);
  printf("{\"client\": \"pqa_client\", \"learners\": %d, \"trainings\": %" PRId64 ", \"questions_asked\": %" PRIu64
         ", \"seconds\": %.3f, \"questions_per_s\": %.1f, \"mean_window_questions_per_s\": %.1f, \"last_window_precision_pct\": %.2f, "
         "\"failed\": %s}\n",
         nLearners, sh.nTrainingsMax, asked, total, (double)asked / total, sh.nReports ? sh.sumRate / (double)sh.nReports : 0.0,
         sh.lastPrecision, sh.failed ? "true" : "false");
  fclose(sh.progress);
  CiReleasePqaEngine(sh.engine);
  return sh.failed ? 1 : 0;
}
