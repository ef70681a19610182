// probqa_b200: the extern "C" layer of libPqaCore.so. Every PqaCInterop.h symbol of the reference
// (PqaCore/Interface/PqaCInterop.h:48-108, definitions in PqaCore/PqaCInterop.cpp) is exported with the same
// signature and error convention; every reference entry point is implemented (Logger_Init / SetLogger are accepted and
// ignored). Combinations this engine does not offer (e.g. maintenance mode on a sharded engine) return the reference's
// NotImplemented error object, its own convention for unimplemented engine features (CudaEngine.cpp:62-98).
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/PqaB200Ext.h"
#include "pqa_engine.h"
#include "pqa_group.h"

#include <cstdlib>
#include <string>

using namespace pqa;

namespace {

struct Factory { int unused; };
Factory g_factory;

inline Engine *E(void *pv) { return static_cast<Engine *>(pv); }

PqaError *NullEngine() { return MakeError(ErrCode::NullArgument, "pvEngine is NULL"); }
// Every entry point that enters an engine passes here first: a NULL handle, or an engine after PqaEngine_Shutdown
// (MaintenanceSwitch.h:138-141: operations on a shut-down engine fail with ObjectShutDown).
PqaError *Gate(void *pvEngine) {
  if (!pvEngine) return NullEngine();
  if (static_cast<Engine *>(pvEngine)->IsShutDown())
    return MakeError(ErrCode::ObjectShutDown, "The engine is shut down.", "MaintenanceSwitch::TryEnterSpecific()");
  return nullptr;
}

// ReturnPqaError (PqaCInterop.cpp:56-61): NULL on success, owned error otherwise.
inline void *Ret(PqaError *e) { return e; }
// AssignPqaError (PqaCInterop.cpp:45-54)
inline void Assign(void **ppError, PqaError *e) {
  if (ppError) *ppError = e;
  else if (e) delete e;
}

template <typename F> PqaError *Guard(F &&f) {
  try { return f(); }
  catch (const CudaFail &cf) { return ErrCuda(cf.code, cf.what(), cf.file, cf.line); }
  catch (const std::exception &ex) { return ErrStd(ex.what()); }
  catch (...) { return MakeError(ErrCode::Internal, "unknown exception"); }
}

PqaError *CreateEngineImpl(const CiEngineDefinition *pEngDef, const CiB200Options *pOpts, Engine **out) {
  *out = nullptr;
  if (!pEngDef) return MakeError(ErrCode::NullArgument, "pEngDef is NULL");
  // PqaEngineBaseFactory::CreateCpuEngine (PqaEngineBaseFactory.cpp:16-27): only Double precision;
  // CheckDimensions (:113-122) with the minimums of PqaEngineBaseFactory.h:15-17.
  if (pEngDef->_precType != 3)
    return ErrNotImplemented("B200 engine on precision type other than double (precType must be 3).");
  if (pEngDef->_nAnswers < 2 || pEngDef->_nQuestions < 1 || pEngDef->_nTargets < 2)
    return ErrInsufficientDims(pEngDef->_nAnswers, pEngDef->_nQuestions, pEngDef->_nTargets);
  CiB200Options opts;
  if (pOpts) opts = *pOpts;
  else { std::memset(&opts, 0, sizeof(opts)); opts._device = -1; }
  return Guard([&]() -> PqaError * { *out = new Engine(*pEngDef, opts); return nullptr; });
}

// PQA_B200_SHARDS=N (with PQA_B200_SHARD_AXIS / _DEVICES / _EXACT / PQA_B200_SHARD_MAX_BATCH): the reference's factory
// entry points hand out a sharded engine group, so that an unmodified client of the reference ABI runs on N GPUs.
bool EnvGroupOptions(CiB200GroupOptions *g) {
  const char *n = std::getenv("PQA_B200_SHARDS");
  if (!n || std::atoi(n) < 2) return false;
  std::memset(g, 0, sizeof(*g));
  g->_nShards = std::atoi(n);
  const char *axis = std::getenv("PQA_B200_SHARD_AXIS");
  g->_axis = (axis && std::string(axis) == "questions") ? 0 : 1;
  for (int r = 0; r < 8; r++) g->_devices[r] = -1;
  if (const char *d = std::getenv("PQA_B200_SHARD_DEVICES")) {
    int r = 0;
    for (const char *p = d; *p && r < 8; r++) {
      g->_devices[r] = std::atoi(p);
      while (*p && *p != ',') p++;
      if (*p == ',') p++;
    }
  }
  const char *ex = std::getenv("PQA_B200_SHARD_EXACT");
  g->_exactOrder = ex && std::atoi(ex) != 0;
  const char *mb = std::getenv("PQA_B200_SHARD_MAX_BATCH");
  g->_maxBatch = mb ? std::atoll(mb) : 0;
  return true;
}

PqaError *CreateGroupImpl(const CiEngineDefinition *pEngDef, const CiB200Options *pOpts, const CiB200GroupOptions *pG, Engine **out) {
  *out = nullptr;
  if (!pEngDef || !pG) return MakeError(ErrCode::NullArgument, "pEngDef / pGroupOpts is NULL");
  if (pEngDef->_precType != 3)
    return ErrNotImplemented("B200 engine on precision type other than double (precType must be 3).");
  if (pEngDef->_nAnswers < 2 || pEngDef->_nQuestions < 1 || pEngDef->_nTargets < 2)
    return ErrInsufficientDims(pEngDef->_nAnswers, pEngDef->_nQuestions, pEngDef->_nTargets);
  CiB200Options opts;
  if (pOpts) opts = *pOpts;
  else { std::memset(&opts, 0, sizeof(opts)); opts._device = -1; }
  return Guard([&]() -> PqaError * { *out = new ShardGroup(*pEngDef, opts, *pG); return nullptr; });
}

} // namespace

extern "C" {

PQACORE_API void CiDebugBreak(void) {}

PQACORE_API uint8_t Logger_Init(void **ppStrErr, const char *baseName) {
  // the reference opens a file logger (PqaCInterop.cpp:145-172); this library logs to stderr only
  (void)baseName;
  if (ppStrErr) *ppStrErr = nullptr;
  return 1;
}
PQACORE_API void CiReleaseString(void *pvString) { delete[] static_cast<char *>(pvString); }

PQACORE_API void *CiGetPqaEngineFactory(void) { return &g_factory; }

PQACORE_API void *PqaEngineFactory_CreateCpuEngine(void *pvFactory, void **ppError, const CiEngineDefinition *pEngDef) {
  (void)pvFactory;
  Engine *eng = nullptr;
  CiB200GroupOptions g;
  if (EnvGroupOptions(&g)) Assign(ppError, CreateGroupImpl(pEngDef, nullptr, &g, &eng));
  else Assign(ppError, CreateEngineImpl(pEngDef, nullptr, &eng));
  return eng;
}
PQACORE_API void *PqaB200_CreateShardedEngine(void **ppError, const CiEngineDefinition *pEngDef, const CiB200Options *pOpts,
                                              const CiB200GroupOptions *pGroupOpts) {
  Engine *eng = nullptr;
  Assign(ppError, CreateGroupImpl(pEngDef, pOpts, pGroupOpts, &eng));
  return eng;
}
PQACORE_API void *PqaB200_LoadShardedEngine(void **ppError, const char *filePath, const CiB200Options *pOpts,
                                            const CiB200GroupOptions *pGroupOpts) {
  if (!pGroupOpts) { Assign(ppError, MakeError(ErrCode::NullArgument, "pGroupOpts is NULL")); return nullptr; }
  CiB200Options opts;
  if (pOpts) opts = *pOpts;
  else { std::memset(&opts, 0, sizeof(opts)); opts._device = -1; }
  PqaError *err = nullptr;
  Engine *eng = nullptr;
  PqaError *g = Guard([&]() -> PqaError * { eng = ShardGroup::LoadKBGroup(filePath, opts, *pGroupOpts, &err); return nullptr; });
  Assign(ppError, g ? g : err);
  return eng;
}
PQACORE_API int32_t PqaB200_GetShardCount(void *pvEngine) {
  if (!pvEngine) return 0;
  const ShardGroup *g = dynamic_cast<ShardGroup *>(E(pvEngine));
  return g ? g->shardCount() : 1;
}
PQACORE_API void *PqaB200_CreateEngine(void **ppError, const CiEngineDefinition *pEngDef, const CiB200Options *pOpts) {
  Engine *eng = nullptr;
  Assign(ppError, CreateEngineImpl(pEngDef, pOpts, &eng));
  return eng;
}
PQACORE_API void *PqaEngineFactory_LoadCpuEngine(void *pvFactory, void **ppError, const char *filePath,
                                                 uint64_t memPoolMaxBytes) {
  (void)pvFactory; (void)memPoolMaxBytes;
  CiB200Options opts; std::memset(&opts, 0, sizeof(opts)); opts._device = -1;
  PqaError *err = nullptr;
  Engine *eng = nullptr;
  CiB200GroupOptions go;
  const bool grouped = EnvGroupOptions(&go);
  PqaError *g = Guard([&]() -> PqaError * {
    eng = grouped ? static_cast<Engine *>(ShardGroup::LoadKBGroup(filePath, opts, go, &err)) : Engine::LoadKB(filePath, opts, &err);
    return nullptr;
  });
  Assign(ppError, g ? g : err);
  return eng;
}

PQACORE_API void *PqaB200_LoadEngine(void **ppError, const char *filePath, const CiB200Options *pOpts) {
  CiB200Options opts;
  if (pOpts) opts = *pOpts;
  else { std::memset(&opts, 0, sizeof(opts)); opts._device = -1; }
  PqaError *err = nullptr;
  Engine *eng = nullptr;
  PqaError *g = Guard([&]() -> PqaError * { eng = Engine::LoadKB(filePath, opts, &err); return nullptr; });
  Assign(ppError, g ? g : err);
  return eng;
}
PQACORE_API void *PqaB200_SaveKBShard(void *pvEngine, const char *filePath, int32_t writeFrame) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->SaveKBShard(filePath, writeFrame != 0); }));
}

PQACORE_API void CiReleasePqaError(void *pvErr) { delete static_cast<PqaError *>(pvErr); }
PQACORE_API void *PqaError_ToString(void *pvError, const uint8_t withParams) {
  const std::string s = pvError ? static_cast<PqaError *>(pvError)->ToString(withParams != 0) : std::string("Success");
  char *out = new char[s.size() + 1];
  std::memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

PQACORE_API void CiReleasePqaEngine(void *pvEngine) { delete E(pvEngine); }

PQACORE_API void *PqaEngine_Train(void *pvEngine, int64_t nQuestions, const CiAnsweredQuestion *const pAQs,
                                  const int64_t iTarget, const double amount) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->Train(nQuestions, pAQs, iTarget, amount); }));
}

// Permanent <-> compact id maps (BaseEngine.cpp:154-218, PermanentIdManager.cpp); ids that do not map give -1.
static uint8_t MapIds(void *pvEngine, int kind, bool permFromComp, int64_t count, int64_t *pIds) {
  if (!pvEngine || E(pvEngine)->IsShutDown() || (count > 0 && !pIds)) return 0;
  return E(pvEngine)->MapIds(kind, permFromComp, count, pIds) ? 1 : 0;
}
PQACORE_API uint8_t PqaEngine_QuestionPermFromComp(void *pvEngine, const int64_t count, int64_t *pIds) { return MapIds(pvEngine, 0, true, count, pIds); }
PQACORE_API uint8_t PqaEngine_QuestionCompFromPerm(void *pvEngine, const int64_t count, int64_t *pIds) { return MapIds(pvEngine, 0, false, count, pIds); }
PQACORE_API uint8_t PqaEngine_TargetPermFromComp(void *pvEngine, const int64_t count, int64_t *pIds) { return MapIds(pvEngine, 1, true, count, pIds); }
PQACORE_API uint8_t PqaEngine_TargetCompFromPerm(void *pvEngine, const int64_t count, int64_t *pIds) { return MapIds(pvEngine, 1, false, count, pIds); }
PQACORE_API uint8_t PqaEngine_QuizPermFromComp(void *pvEngine, const int64_t count, int64_t *pIds) { return MapIds(pvEngine, 2, true, count, pIds); }
PQACORE_API uint8_t PqaEngine_QuizCompFromPerm(void *pvEngine, const int64_t count, int64_t *pIds) { return MapIds(pvEngine, 2, false, count, pIds); }
PQACORE_API uint8_t PqaEngine_EnsurePermQuizGreater(void *pvEngine, const int64_t bound) {
  return pvEngine && !E(pvEngine)->IsShutDown() && E(pvEngine)->EnsurePermQuizGreater(bound) ? 1 : 0;
}
PQACORE_API uint8_t PqaEngine_RemapQuizPermId(void *pvEngine, const int64_t srcPermId, const int64_t destPermId) {
  return pvEngine && !E(pvEngine)->IsShutDown() && E(pvEngine)->RemapQuizPermId(srcPermId, destPermId) ? 1 : 0;
}

PQACORE_API uint64_t PqaEngine_GetTotalQuestionsAsked(void *pvEngine, void **ppError) {
  if (PqaError *g_ = Gate(pvEngine)) { Assign(ppError, g_); return 0; }
  Assign(ppError, nullptr);
  return E(pvEngine)->GetTotalQuestionsAsked();
}
PQACORE_API uint8_t PqaEngine_CopyDims(void *pvEngine, CiEngineDimensions *pDims) {
  if (!pvEngine || !pDims) return 0;
  *pDims = E(pvEngine)->IsShutDown() ? CiEngineDimensions{0, 0, 0} : E(pvEngine)->CopyDims();   // BaseEngine.cpp:315
  return 1;
}
PQACORE_API int64_t PqaEngine_StartQuiz(void *pvEngine, void **ppError) {
  if (PqaError *g_ = Gate(pvEngine)) { Assign(ppError, g_); return -1; }
  PqaError *err = nullptr;
  int64_t id = -1;
  PqaError *g = Guard([&]() -> PqaError * { id = E(pvEngine)->StartQuiz(&err); return nullptr; });
  if (g) { delete err; err = g; id = -1; }
  Assign(ppError, err);
  return id;
}
PQACORE_API int64_t PqaEngine_ResumeQuiz(void *pvEngine, void **ppError, const int64_t nAnswered,
                                         const CiAnsweredQuestion *const pAQs) {
  if (PqaError *g_ = Gate(pvEngine)) { Assign(ppError, g_); return -1; }
  PqaError *err = nullptr;
  int64_t id = -1;
  PqaError *g = Guard([&]() -> PqaError * { id = E(pvEngine)->ResumeQuiz(&err, nAnswered, pAQs); return nullptr; });
  if (g) { delete err; err = g; id = -1; }
  Assign(ppError, err);
  return id;
}
PQACORE_API int64_t PqaEngine_NextQuestion(void *pvEngine, void **ppError, const int64_t iQuiz) {
  if (PqaError *g_ = Gate(pvEngine)) { Assign(ppError, g_); return -1; }
  PqaError *err = nullptr;
  int64_t q = -1;
  PqaError *g = Guard([&]() -> PqaError * { q = E(pvEngine)->NextQuestion(&err, iQuiz); return nullptr; });
  if (g) { delete err; err = g; q = -1; }
  Assign(ppError, err);
  return q;
}
PQACORE_API void *PqaEngine_RecordAnswer(void *pvEngine, const int64_t iQuiz, const int64_t iAnswer) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->RecordAnswer(iQuiz, iAnswer); }));
}
PQACORE_API void *PqaEngine_ClearOldQuizzes(void *pvEngine, const int64_t maxCount, const double maxAgeSec) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ClearOldQuizzes(maxCount, maxAgeSec); }));
}
PQACORE_API int64_t PqaEngine_GetActiveQuestionId(void *pvEngine, void **ppError, const int64_t iQuiz) {
  if (PqaError *g_ = Gate(pvEngine)) { Assign(ppError, g_); return -1; }
  PqaError *err = nullptr;
  int64_t q = -1;
  PqaError *g = Guard([&]() -> PqaError * { q = E(pvEngine)->GetActiveQuestionId(&err, iQuiz); return nullptr; });
  if (g) { delete err; err = g; q = -1; }
  Assign(ppError, err);
  return q;
}
PQACORE_API void *PqaEngine_SetActiveQuestion(void *pvEngine, const int64_t iQuiz, const int64_t iQuestion) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->SetActiveQuestion(iQuiz, iQuestion); }));
}
PQACORE_API int64_t PqaEngine_ListTopTargets(void *pvEngine, void **ppError, const int64_t iQuiz,
                                             const int64_t maxCount, CiRatedTarget *pDest) {
  if (PqaError *g_ = Gate(pvEngine)) { Assign(ppError, g_); return -1; }
  PqaError *err = nullptr;
  int64_t n = -1;
  PqaError *g = Guard([&]() -> PqaError * { n = E(pvEngine)->ListTopTargets(&err, iQuiz, maxCount, pDest); return nullptr; });
  if (g) { delete err; err = g; n = -1; }
  Assign(ppError, err);
  return n;
}
PQACORE_API void *PqaEngine_RecordQuizTarget(void *pvEngine, const int64_t iQuiz, const int64_t iTarget,
                                             const double amount) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->RecordQuizTarget(iQuiz, iTarget, amount); }));
}
PQACORE_API void *PqaEngine_ReleaseQuiz(void *pvEngine, const int64_t iQuiz) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ReleaseQuiz(iQuiz); }));
}
PQACORE_API void *PqaEngine_SaveKB(void *pvEngine, const char *const filePath, const uint8_t bDoubleBuffer) {
  (void)bDoubleBuffer;
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->SaveKB(filePath); }));
}

PQACORE_API void *PqaEngine_StartMaintenance(void *pvEngine, const bool forceQuizzes) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->StartMaintenance(forceQuizzes); }));
}
PQACORE_API void *PqaEngine_FinishMaintenance(void *pvEngine) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->FinishMaintenance(); }));
}
PQACORE_API void *PqaEngine_AddQsTs(void *pvEngine, const int64_t nQuestions, CiAddQorTParam *pAddQuestionParams,
                                    const int64_t nTargets, CiAddQorTParam *pAddTargetParams) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->AddQsTs(nQuestions, pAddQuestionParams, nTargets, pAddTargetParams); }));
}
PQACORE_API void *PqaEngine_RemoveQuestions(void *pvEngine, const int64_t nQuestions, const int64_t *pQIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->RemoveQuestions(nQuestions, pQIds); }));
}
PQACORE_API void *PqaEngine_RemoveTargets(void *pvEngine, const int64_t nTargets, const int64_t *pTIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->RemoveTargets(nTargets, pTIds); }));
}
PQACORE_API void *PqaEngine_Compact(void *pvEngine, int64_t *pnQuestions, int64_t const **const ppOldQuestions,
                                    int64_t *pnTargets, int64_t const **const ppOldTargets) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->Compact(pnQuestions, ppOldQuestions, pnTargets, ppOldTargets); }));
}
PQACORE_API void CiReleaseCompaction(const int64_t *p) { delete[] p; }
PQACORE_API void *PqaEngine_Shutdown(void *pvEngine, const char *const saveFilePath) {
  if (!pvEngine) return NullEngine();
  return Ret(Guard([&] { return E(pvEngine)->Shutdown(saveFilePath); }));      // a second Shutdown answers ObjectShutDown itself
}
PQACORE_API void *PqaEngine_SetLogger(void *pvEngine, void *pSRLogger) {
  (void)pSRLogger;
  return pvEngine ? nullptr : NullEngine();
}

// ------------------------------------------------------------------ PqaB200Ext.h
PQACORE_API int32_t PqaB200_GetEmulatedWorkers(void *pvEngine) { return pvEngine ? E(pvEngine)->emulatedWorkers() : -1; }
PQACORE_API int32_t PqaB200_GetDevice(void *pvEngine) { return pvEngine ? E(pvEngine)->device() : -1; }
PQACORE_API const char *PqaB200_BuildInfo(void) {
  return "probqa_b200 libPqaCore: sm_100a, fp64, -fmad=false, built " __DATE__ " " __TIME__;
}
PQACORE_API void *PqaB200_HostLogicSelfTest(void) {
  const std::string m = HostLogicSelfTest();
  if (m.empty()) return nullptr;
  char *out = new char[m.size() + 1];
  std::memcpy(out, m.c_str(), m.size() + 1);
  return out;
}
PQACORE_API void *PqaEngine_CopyATargets(void *pvEngine, int64_t iQuestion, int64_t iAnswer, int64_t maxTargets, double *pFreqs) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->CopyATargets(iQuestion, iAnswer, maxTargets, pFreqs); }));
}
PQACORE_API void *PqaEngine_CopyDTargets(void *pvEngine, int64_t iQuestion, int64_t maxTargets, double *pFreqs) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->CopyDTargets(iQuestion, maxTargets, pFreqs); }));
}
PQACORE_API void *PqaEngine_CopyBTargets(void *pvEngine, int64_t maxTargets, double *pFreqs) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->CopyBTargets(maxTargets, pFreqs); }));
}
PQACORE_API void *PqaB200_UploadKB(void *pvEngine, const double *sA, const double *mD, const double *vB) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->UploadKB(sA, mD, vB); }));
}
PQACORE_API void *PqaB200_DownloadKB(void *pvEngine, double *sA, double *mD, double *vB) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->DownloadKB(sA, mD, vB); }));
}
PQACORE_API void *PqaEngine_StartQuizBatch(void *pvEngine, int64_t n, int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->StartQuizBatch(n, pQuizIds); }));
}
PQACORE_API void *PqaEngine_ResumeQuizBatch(void *pvEngine, int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs,
                                            int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ResumeQuizBatch(n, pCounts, pAQs, pQuizIds); }));
}
PQACORE_API void *PqaEngine_NextQuestionBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds,
                                              const uint64_t *pRandoms, int64_t *pQuestions, void **ppErrors) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->NextQuestionBatch(n, pQuizIds, pRandoms, pQuestions, ppErrors); }));
}
PQACORE_API void *PqaEngine_RecordAnswerBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->RecordAnswerBatch(n, pQuizIds, pAnswers); }));
}
PQACORE_API void *PqaEngine_SetActiveQuestionBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pQuestions) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->SetActiveQuestionBatch(n, pQuizIds, pQuestions); }));
}
PQACORE_API void *PqaEngine_ListTopTargetsBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, int64_t maxCount,
                                                CiRatedTarget *pDest, int64_t *pCounts) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ListTopTargetsBatch(n, pQuizIds, maxCount, pDest, pCounts); }));
}
PQACORE_API void *PqaEngine_RecordQuizTargetBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds,
                                                  const int64_t *pTargets, const double *pAmounts) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->RecordQuizTargetBatch(n, pQuizIds, pTargets, pAmounts); }));
}
PQACORE_API void *PqaEngine_ReleaseQuizBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ReleaseQuizBatch(n, pQuizIds); }));
}
PQACORE_API void *PqaB200_CopyQuizPriors(void *pvEngine, int64_t iQuiz, double *pPriors) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->CopyQuizPriors(iQuiz, pPriors); }));
}
PQACORE_API void *PqaB200_SetQuizPriors(void *pvEngine, int64_t iQuiz, const double *pPriors) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->SetQuizPriors(iQuiz, pPriors); }));
}
PQACORE_API void *PqaB200_EvalQuestions(void *pvEngine, int64_t n, const int64_t *pQuizIds, double *pPriorities,
                                        double *pRunLength, double *pGrandTotals, int64_t *pnChunks) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->EvalQuestions(n, pQuizIds, pPriorities, pRunLength, pGrandTotals, pnChunks); }));
}
PQACORE_API void *PqaB200_EvalQuestionsDetailed(void *pvEngine, int64_t iQuiz, double *pW, double *pH, double *pV,
                                                double *pLack, double *pPriorities) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->EvalQuestionsDetailed(iQuiz, pW, pH, pV, pLack, pPriorities); }));
}
PQACORE_API void *PqaB200_EvalQuestionsDetailedBatch(void *pvEngine, int64_t n, const int64_t *pQuizIds, double *pW, double *pH,
                                                     double *pV, double *pLack, double *pPriorities) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->EvalQuestionsDetailedBatch(n, pQuizIds, pW, pH, pV, pLack, pPriorities); }));
}
PQACORE_API void *PqaB200_SetEvalKernel(void *pvEngine, int32_t which) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->SetEvalKernel(which, 0, 0, 0); }));
}
PQACORE_API void *PqaB200_SetEvalTuning(void *pvEngine, int32_t which, int64_t chunkTargets, int64_t quizzesPerCta,
                                        int32_t kahanLanesPerThread) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->SetEvalKernel(which, chunkTargets, quizzesPerCta, kahanLanesPerThread); }));
}
PQACORE_API void *PqaB200_ShardEval(void *pvEngine, int64_t n, const int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ShardEval(n, pQuizIds); }));
}
PQACORE_API void *PqaB200_ShardSelect(void *pvEngine, int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms,
                                      int64_t *pQuestions, void **ppErrors) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ShardSelect(n, pQuizIds, pRandoms, pQuestions, ppErrors); }));
}
PQACORE_API void *PqaB200_ShardRecordAnswerBegin(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ShardRecordAnswerBegin(n, pQuizIds, pAnswers); }));
}
PQACORE_API void *PqaB200_ShardRecordAnswerEnd(void *pvEngine, int64_t n, const int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ShardRecordAnswerEnd(n, pQuizIds); }));
}
PQACORE_API void *PqaB200_ShardBuffer(void *pvEngine, int32_t which, void **ppDevice, int64_t *pCount) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ShardBuffer(which, ppDevice, pCount); }));
}
PQACORE_API void *PqaB200_GetQuestionShard(void *pvEngine, int64_t *pFirst, int64_t *pCount) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  if (pFirst) *pFirst = E(pvEngine)->questionShardFirst();
  if (pCount) *pCount = E(pvEngine)->questionShardCount();
  return nullptr;
}
PQACORE_API void *PqaB200_TShardEvalW(void *pvEngine, int64_t n, const int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->TShardEvalW(n, pQuizIds); }));
}
PQACORE_API void *PqaB200_TShardEvalHVL(void *pvEngine, int64_t n, const int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->TShardEvalHVL(n, pQuizIds); }));
}
PQACORE_API void *PqaB200_TShardPriority(void *pvEngine, int64_t n, const int64_t *pQuizIds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->TShardPriority(n, pQuizIds); }));
}
PQACORE_API void *PqaB200_GetTargetShard(void *pvEngine, int64_t *pFirst, int64_t *pCount) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  if (pFirst) *pFirst = E(pvEngine)->targetShardFirst();
  if (pCount) *pCount = E(pvEngine)->targetShardCount();
  return nullptr;
}
PQACORE_API void *PqaB200_FillBinarySearchKB(void *pvEngine, double rounds) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->FillBinarySearchKB(rounds); }));
}
PQACORE_API void *PqaB200_P2PInit(void *pvEngine, int32_t rank, int32_t nRanks, int64_t maxQuizzes, void **ppBase, int64_t *pBytes) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PInit(rank, nRanks, maxQuizzes, ppBase, pBytes); }));
}
PQACORE_API void *PqaB200_P2PExportHandle(void *pvEngine, uint8_t *pHandle64) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PExportHandle(pHandle64); }));
}
PQACORE_API void *PqaB200_P2POpenHandle(void *pvEngine, const uint8_t *pHandle64, void **ppPeerBase) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2POpenHandle(pHandle64, ppPeerBase); }));
}
PQACORE_API void *PqaB200_P2PConnect(void *pvEngine, void *const *pBases) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PConnect(pBases); }));
}
PQACORE_API void *PqaB200_P2PNextQuestionBegin(void *pvEngine, int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PNextQuestionBegin(n, pQuizIds, pRandoms); }));
}
PQACORE_API void *PqaB200_P2PNextQuestionEnd(void *pvEngine, int64_t n, const int64_t *pQuizIds, int64_t *pQuestions, void **ppErrors) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PNextQuestionEnd(n, pQuizIds, pQuestions, ppErrors); }));
}
PQACORE_API void *PqaB200_P2PRecordAnswerBegin(void *pvEngine, int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PRecordAnswerBegin(n, pQuizIds, pAnswers); }));
}
PQACORE_API void *PqaB200_P2PRecordAnswerEnd(void *pvEngine) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PRecordAnswerEnd(); }));
}
PQACORE_API void *PqaB200_P2PLastPhaseMs(void *pvEngine, double *pMs5) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  if (!pMs5) return Ret(MakeError(ErrCode::NullArgument, "pMs5"));
  return Ret(Guard([&] { return E(pvEngine)->P2PLastPhaseMs(pMs5); }));
}
PQACORE_API void *PqaB200_AnomalyCounts(void *pvEngine, uint64_t *pCounts3) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  if (!pCounts3) return Ret(MakeError(ErrCode::NullArgument, "pCounts3"));
  return Ret(Guard([&] { return E(pvEngine)->AnomalyCounts(pCounts3); }));
}
PQACORE_API void *PqaB200_P2PSetExactOrder(void *pvEngine, int32_t on) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->P2PSetExactOrder(on); }));
}
PQACORE_API void *PqaB200_ResidentBind(void *pvEngine, int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ResidentBind(n, pQuizIds, pRandoms); }));
}
PQACORE_API void *PqaB200_ResidentStep(void *pvEngine) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ResidentStep(); }));
}
PQACORE_API void *PqaB200_ResidentFetch(void *pvEngine, int64_t *pQuestions) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->ResidentFetch(pQuestions); }));
}
PQACORE_API double PqaB200_ResidentLastEvalMs(void *pvEngine) { return pvEngine ? E(pvEngine)->ResidentLastEvalMs() : -1.0; }
PQACORE_API void *PqaB200_Synchronize(void *pvEngine) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->Synchronize(); }));
}
PQACORE_API void *PqaB200_EventCreate(void) {
  cudaEvent_t ev = nullptr;
  if (cudaEventCreate(&ev) != cudaSuccess) return nullptr;
  return ev;
}
PQACORE_API void PqaB200_EventDestroy(void *pvEvent) { if (pvEvent) cudaEventDestroy((cudaEvent_t)pvEvent); }
PQACORE_API void *PqaB200_EventRecord(void *pvEngine, void *pvEvent) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  const cudaError_t e = cudaEventRecord((cudaEvent_t)pvEvent, E(pvEngine)->stream());
  return e == cudaSuccess ? nullptr : ErrCuda((int)e, "cudaEventRecord", __FILE__, __LINE__);
}
PQACORE_API void *PqaB200_EventSynchronize(void *pvEvent) {
  const cudaError_t e = cudaEventSynchronize((cudaEvent_t)pvEvent);
  return e == cudaSuccess ? nullptr : ErrCuda((int)e, "cudaEventSynchronize", __FILE__, __LINE__);
}
PQACORE_API double PqaB200_EventElapsedMs(void *pvStart, void *pvStop) {
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, (cudaEvent_t)pvStart, (cudaEvent_t)pvStop) != cudaSuccess) return -1.0;
  return (double)ms;
}
PQACORE_API uint64_t PqaB200_KernelLaunchCount(void *pvEngine) { (void)pvEngine; return kernel_launch_count(); }
PQACORE_API void *PqaB200_FlushL2(void *pvEngine) {
  if (PqaError *g_ = Gate(pvEngine)) return g_;
  return Ret(Guard([&] { return E(pvEngine)->FlushL2(); }));
}

} // extern "C"
