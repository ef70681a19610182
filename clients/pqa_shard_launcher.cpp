// probqa_b200: one process per GPU without Python or torch -- a C++ launcher for the sharded engines of a box.
//
// The parent forks one child per GPU *before* touching CUDA. Child r creates the shard engine of rank r through the C ABI
// (PqaB200_CreateEngine with a question or target shard on device r), fills its part of the synthetic KB on the device,
// and joins the peer-memory exchange: PqaB200_P2PInit -> the 64-byte cudaIpc handle of its inbox goes to the parent over
// a UNIX socket pair, the parent hands every child the full list, PqaB200_P2POpenHandle / P2PConnect. From then on the
// children issue the same PqaB200_P2PNextQuestionBegin/End (and RecordAnswer) calls in lockstep; the kernels exchange over
// NVLink and no host data moves between the processes. The parent is only a rendezvous (handle exchange, barriers, the
// final report). Workload = bench.py's sharded leg (BASELINE config 4 by default): B quizzes at depths 0/3/8, `--steps`
// timed NextQuestion batches; one JSON line on stdout.
//
//   pqa_shard_launcher --gpus 8 --axis targets --exact-order --questions 10000 --answers 5 --targets 100000 --batch 64
//
// Only include/PqaCInterop.h + include/PqaB200Ext.h are used (link with -lPqaCore).
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/PqaB200Ext.h"

namespace {

struct Args {
  int gpus = 2, steps = 10, warmup = 3, exact = 0;
  std::string axis = "targets";
  int64_t Q = 10000, K = 5, T = 100000, B = 64;
};

std::string ErrText(void *err) {
  if (!err) return "";
  void *s = PqaError_ToString(err, 1);
  std::string out = s ? static_cast<const char *>(s) : "(no text)";
  CiReleaseString(s);
  CiReleasePqaError(err);
  return out;
}
#define CHECK(call)                                                                        \
  do {                                                                                     \
    if (void *e_ = (call)) { fprintf(stderr, "[rank %d] %s: %s\n", rank, #call, ErrText(e_).c_str()); return 2; } \
  } while (0)

bool WriteAll(int fd, const void *p, size_t n) {
  const char *c = static_cast<const char *>(p);
  while (n > 0) { const ssize_t w = write(fd, c, n); if (w <= 0) return false; c += w; n -= (size_t)w; }
  return true;
}
bool ReadAll(int fd, void *p, size_t n) {
  char *c = static_cast<char *>(p);
  while (n > 0) { const ssize_t r = read(fd, c, n); if (r <= 0) return false; c += r; n -= (size_t)r; }
  return true;
}
// child side of the rendezvous: send `out` (nOut bytes), receive `in` (nIn bytes)
bool Exchange(int fd, const void *out, size_t nOut, void *in, size_t nIn) { return WriteAll(fd, out, nOut) && ReadAll(fd, in, nIn); }
bool Barrier(int fd) { char c = 1; return Exchange(fd, &c, 1, &c, 1); }

// probqa_b200/synth.py: hidden_target, quiz_prefix (questions from a fixed LCG without repeats, answers by the
// binary-search rule of DichotomyTest.cpp:50-67)
int64_t HiddenTarget(int64_t b, int64_t T) { return (int64_t)(((unsigned __int128)b * 2654435761ull) % (uint64_t)T); }
std::vector<std::pair<int64_t, int64_t>> QuizPrefix(int64_t b, int depth, int64_t Q, int64_t T, int64_t K) {
  const int64_t t = HiddenTarget(b, T), w = std::max<int64_t>(1, (32 * T) / 1000);
  std::vector<std::pair<int64_t, int64_t>> out;
  uint64_t x = (1103515245ull * (uint64_t)(b + 12345) + 12345ull) & 0x7FFFFFFFull;
  while ((int)out.size() < depth) {
    x = (1103515245ull * x + 12345ull) & 0x7FFFFFFFull;
    const int64_t q = (int64_t)(x % (uint64_t)Q);
    bool seen = false;
    for (auto &p : out) seen |= p.first == q;
    if (seen) continue;
    const int64_t piv = (q * T) / Q;
    int64_t a = t < piv - w ? 0 : t < piv ? 1 : t == piv ? 2 : t <= piv + w ? 3 : 4;
    out.emplace_back(q, std::min(a, K - 1));
  }
  return out;
}

// CalcSplit (SRPoolRunner.h:96-110) over questions, or over 4-target vectors for target shards (probqa_b200/sharded.py)
void ShardRange(int64_t units, int parts, int p, int64_t *first, int64_t *count) {
  const int64_t quot = units / parts, rem = units % parts;
  *first = p * quot + std::min<int64_t>(p, rem);
  *count = quot + (p < rem ? 1 : 0);
}

int Child(const Args &a, int rank, int fd) {
  CiEngineDefinition def;
  memset(&def, 0, sizeof(def));
  def._nAnswers = a.K; def._nQuestions = a.Q; def._nTargets = a.T;
  def._precType = 3; def._initAmount = 0.1; def._memPoolMaxBytes = 0;
  CiB200Options opt;
  memset(&opt, 0, sizeof(opt));
  opt._device = rank; opt._rngSeed = 1234; opt._initialQuizCapacity = a.B;
  int64_t first, count;
  if (a.axis == "questions") {
    ShardRange(a.Q, a.gpus, rank, &first, &count);
    opt._questionShardFirst = first; opt._questionShardCount = a.gpus > 1 ? count : 0;
  } else {
    ShardRange((a.T + 3) / 4, a.gpus, rank, &first, &count);
    opt._targetShardFirst = 4 * first; opt._targetShardCount = a.gpus > 1 ? std::min<int64_t>(4 * count, a.T - 4 * first) : 0;
  }
  void *err = nullptr;
  void *eng = PqaB200_CreateEngine(&err, &def, &opt);
  if (!eng) { fprintf(stderr, "[rank %d] PqaB200_CreateEngine: %s\n", rank, ErrText(err).c_str()); return 2; }
  CHECK(PqaB200_FillBinarySearchKB(eng, 3.0));
  if (a.gpus > 1) {
    void *base = nullptr; int64_t bytes = 0;
    CHECK(PqaB200_P2PInit(eng, rank, a.gpus, a.B, &base, &bytes));
    uint8_t mine[64];
    CHECK(PqaB200_P2PExportHandle(eng, mine));
    std::vector<uint8_t> all((size_t)a.gpus * 64);
    if (!Exchange(fd, mine, 64, all.data(), all.size())) { fprintf(stderr, "[rank %d] handle exchange failed\n", rank); return 2; }
    std::vector<void *> bases((size_t)a.gpus, nullptr);
    for (int r = 0; r < a.gpus; r++) {
      if (r == rank) { bases[(size_t)r] = base; continue; }
      CHECK(PqaB200_P2POpenHandle(eng, all.data() + (size_t)r * 64, &bases[(size_t)r]));
    }
    CHECK(PqaB200_P2PConnect(eng, bases.data()));
    if (a.exact && a.axis == "targets") CHECK(PqaB200_P2PSetExactOrder(eng, 1));
    if (!Barrier(fd)) return 2;
  }
  // the batch: quiz b at depth {0, 3, 8}[b % 3]
  std::vector<int64_t> ids((size_t)a.B);
  CHECK(PqaEngine_StartQuizBatch(eng, a.B, ids.data()));
  static const int kDepths[3] = {0, 3, 8};
  std::vector<std::vector<std::pair<int64_t, int64_t>>> states;
  int64_t qevals = 0;
  for (int64_t b = 0; b < a.B; b++) {
    states.push_back(QuizPrefix(b, kDepths[b % 3], a.Q, a.T, a.K));
    qevals += a.Q - (int64_t)states.back().size();
  }
  for (int s = 0; s < 8; s++) {
    std::vector<int64_t> sel, qs, as;
    for (int64_t b = 0; b < a.B; b++)
      if ((int)states[(size_t)b].size() > s) { sel.push_back(ids[(size_t)b]); qs.push_back(states[(size_t)b][(size_t)s].first); as.push_back(states[(size_t)b][(size_t)s].second); }
    if (sel.empty()) break;
    CHECK(PqaEngine_SetActiveQuestionBatch(eng, (int64_t)sel.size(), sel.data(), qs.data()));
    if (a.gpus > 1) {
      CHECK(PqaB200_P2PRecordAnswerBegin(eng, (int64_t)sel.size(), sel.data(), as.data()));
      CHECK(PqaB200_P2PRecordAnswerEnd(eng));
    } else {
      CHECK(PqaEngine_RecordAnswerBatch(eng, (int64_t)sel.size(), sel.data(), as.data()));
    }
  }
  std::vector<uint64_t> randoms((size_t)a.B);
  uint64_t z = 0x9E3779B97F4A7C15ull;
  for (auto &r : randoms) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; r = z; }     // the same draws on every rank
  std::vector<int64_t> chosen((size_t)a.B, -1);
  auto step = [&]() -> void * {
    if (a.gpus > 1) {
      if (void *e = PqaB200_P2PNextQuestionBegin(eng, a.B, ids.data(), randoms.data())) return e;
      return PqaB200_P2PNextQuestionEnd(eng, a.B, ids.data(), chosen.data(), nullptr);
    }
    return PqaEngine_NextQuestionBatch(eng, a.B, ids.data(), randoms.data(), chosen.data(), nullptr);
  };
  for (int s = 0; s < a.warmup; s++) CHECK(step());
  if (!Barrier(fd)) return 2;
  const auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < a.steps; s++) CHECK(step());
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double phases[5] = {0, 0, 0, 0, 0};
  if (a.gpus > 1 && a.axis == "targets") CHECK(PqaB200_P2PLastPhaseMs(eng, phases));
  int64_t checksum = 0;
  for (int64_t b = 0; b < a.B; b++) checksum = (checksum + chosen[(size_t)b] * (b + 1)) % 1000000007;
  struct { double secs; int64_t checksum, qevals; double phases[5]; uint64_t launches; } rep;
  rep.secs = secs; rep.checksum = checksum; rep.qevals = qevals; memcpy(rep.phases, phases, sizeof(phases));
  rep.launches = PqaB200_KernelLaunchCount(eng);
  char ack;
  if (!Exchange(fd, &rep, sizeof(rep), &ack, 1)) return 2;
  CiReleasePqaEngine(eng);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  Args a;
  for (int i = 1; i < argc; i++) {
    const std::string k = argv[i];
    auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : "0"; };
    if (k == "--gpus") a.gpus = atoi(next());
    else if (k == "--steps") a.steps = atoi(next());
    else if (k == "--warmup") a.warmup = atoi(next());
    else if (k == "--axis") a.axis = next();
    else if (k == "--exact-order") a.exact = 1;
    else if (k == "--questions") a.Q = atoll(next());
    else if (k == "--answers") a.K = atoll(next());
    else if (k == "--targets") a.T = atoll(next());
    else if (k == "--batch") a.B = atoll(next());
    else { fprintf(stderr, "usage: %s [--gpus N] [--axis targets|questions] [--exact-order] [--questions Q] [--answers K] [--targets T] [--batch B] [--steps S] [--warmup W]\n", argv[0]); return 1; }
  }
  if (a.gpus < 1 || a.gpus > 8 || (a.axis != "targets" && a.axis != "questions")) { fprintf(stderr, "bad arguments\n"); return 1; }
  std::vector<int> fds((size_t)a.gpus);
  std::vector<pid_t> pids((size_t)a.gpus);
  for (int r = 0; r < a.gpus; r++) {
    int sp[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, sp) != 0) { perror("socketpair"); return 1; }
    const pid_t pid = fork();
    if (pid < 0) { perror("fork"); return 1; }
    if (pid == 0) {                      // child: the parent has not touched CUDA, so this process initialises it afresh
      close(sp[0]);
      for (int q = 0; q < r; q++) close(fds[(size_t)q]);
      _exit(Child(a, r, sp[1]));
    }
    close(sp[1]);
    fds[(size_t)r] = sp[0]; pids[(size_t)r] = pid;
  }
  bool ok = true;
  if (a.gpus > 1) {
    std::vector<uint8_t> all((size_t)a.gpus * 64);
    for (int r = 0; r < a.gpus && ok; r++) ok = ReadAll(fds[(size_t)r], all.data() + (size_t)r * 64, 64);
    for (int r = 0; r < a.gpus && ok; r++) ok = WriteAll(fds[(size_t)r], all.data(), all.size());
  }
  auto barrier = [&]() {
    char c;
    for (int r = 0; r < a.gpus && ok; r++) ok = ReadAll(fds[(size_t)r], &c, 1);
    for (int r = 0; r < a.gpus && ok; r++) ok = WriteAll(fds[(size_t)r], &c, 1);
  };
  if (a.gpus > 1) barrier();             // everybody connected
  barrier();                             // warm-up done: the timed steps start together
  struct Rep { double secs; int64_t checksum, qevals; double phases[5]; uint64_t launches; };
  std::vector<Rep> reps((size_t)a.gpus);
  for (int r = 0; r < a.gpus && ok; r++) ok = ReadAll(fds[(size_t)r], &reps[(size_t)r], sizeof(Rep));
  char ack = 1;
  for (int r = 0; r < a.gpus && ok; r++) ok = WriteAll(fds[(size_t)r], &ack, 1);
  int bad = ok ? 0 : 1;
  for (int r = 0; r < a.gpus; r++) { int st = 0; waitpid(pids[(size_t)r], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1; }
  if (bad) { fprintf(stderr, "pqa_shard_launcher: a rank failed\n"); return 2; }
  double secs = 0, ph[5] = {0, 0, 0, 0, 0};
  bool same = true;
  for (int r = 0; r < a.gpus; r++) {
    secs = std::max(secs, reps[(size_t)r].secs);
    same &= reps[(size_t)r].checksum == reps[0].checksum;
    for (int x = 0; x < 5; x++) ph[x] = std::max(ph[x], reps[(size_t)r].phases[x]);
  }
  printf("{\"launcher\": \"pqa_shard_launcher (C++, one process per GPU, cudaIpc inboxes, no Python/torch)\", \"n_gpus\": %d, \"axis\": \"%s\", "
         "\"exact_order_pipeline\": %s, \"Q\": %" PRId64 ", \"A\": %" PRId64 ", \"T\": %" PRId64 ", \"batch_total\": %" PRId64 ", \"steps\": %d, \"warmup\": %d, "
         "\"ms_per_step\": %.4f, \"value\": %.6g, \"unit\": \"questions/s\", \"all_ranks_chose_the_same_questions\": %s, \"chosen_checksum\": %" PRId64 ", "
         "\"phases_ms\": {\"phase1_W_partials\": %.4f, \"barrier1\": %.4f, \"phase2_HVL_partials\": %.4f, \"barrier2\": %.4f, \"epilogue_select\": %.4f}, "
         "\"timing\": \"wall clock around the P2PNextQuestionBegin/End pairs (host buffers), max over ranks\"}\n",
         a.gpus, a.axis.c_str(), (a.exact && a.axis == "targets") ? "true" : "false", a.Q, a.K, a.T, a.B, a.steps, a.warmup,
         1e3 * secs / a.steps, (double)reps[0].qevals * a.steps / secs, same ? "true" : "false", reps[0].checksum, ph[0], ph[1], ph[2], ph[3], ph[4]);
  return same ? 0 : 3;
}
