// probqa_b200: one engine over several GPUs of a box, in ONE process, behind the reference's C ABI.
// A ShardGroup owns N shard engines (question shards or target shards, pqa_engine.h), one per device, wires their
// peer-memory inboxes together (PqaB200_P2P*: the kernels exchange over NVLink, no host round trip, no NCCL) and
// presents the IPqaEngine method set (PqaCore/Interface/IPqaEngine.h:13-114; validation order and error codes of
// BaseEngine.cpp:399-603 through the inherited shell): StartQuiz / NextQuestion / RecordAnswer / ListTopTargets / RecordQuizTarget / Train /
// ReleaseQuiz / SaveKB ... and their batch forms. It is itself an Engine *shell*: the quiz registry, id validation, error
// objects and the combiner of concurrent one-quiz calls are the base class's; only the device work is fanned out.
#pragma once
#include <memory>
#include <vector>

#include "pqa_engine.h"

namespace pqa {

class ShardGroup : public Engine {
 public:
  ShardGroup(const CiEngineDefinition &def, const CiB200Options &opts, const CiB200GroupOptions &gopts);
  ~ShardGroup() override;
  static ShardGroup *LoadKBGroup(const char *filePath, const CiB200Options &opts, const CiB200GroupOptions &gopts, PqaError **err);

  PqaError *Train(int64_t nQuestions, const CiAnsweredQuestion *pAQs, int64_t iTarget, double amount) override;
  PqaError *StartQuizBatch(int64_t n, int64_t *pQuizIds) override;
  PqaError *NextQuestionBatch(int64_t n, const int64_t *pQuizIds, const uint64_t *pRandoms, int64_t *pQuestions,
                              void **ppErrors) override;
  PqaError *RecordAnswerBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pAnswers) override;
  PqaError *SetActiveQuestionBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pQuestions) override;
  PqaError *ListTopTargetsBatch(int64_t n, const int64_t *pQuizIds, int64_t maxCount, CiRatedTarget *pDest,
                                int64_t *pCounts) override;
  PqaError *RecordQuizTargetBatch(int64_t n, const int64_t *pQuizIds, const int64_t *pTargets, const double *pAmounts) override;
  PqaError *ReleaseQuizBatch(int64_t n, const int64_t *pQuizIds) override;
  PqaError *CopyATargets(int64_t iQuestion, int64_t iAnswer, int64_t maxTargets, double *pFreqs) override;
  PqaError *CopyDTargets(int64_t iQuestion, int64_t maxTargets, double *pFreqs) override;
  PqaError *CopyBTargets(int64_t maxTargets, double *pFreqs) override;
  PqaError *SaveKB(const char *filePath) override;
  PqaError *Shutdown(const char *saveFilePath) override;
  PqaError *UploadKB(const double *sA, const double *mD, const double *vB) override;
  PqaError *DownloadKB(double *sA, double *mD, double *vB) override;
  PqaError *CopyQuizPriors(int64_t iQuiz, double *pPriors) override;
  PqaError *FillBinarySearchKB(double rounds) override;
  PqaError *Synchronize() override;
  PqaError *SetEvalKernel(int32_t which, int64_t chunkTargets, int64_t quizzesPerCta, int32_t kahanLanesPerThread) override;

  PqaError *ResumeQuizBatch(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, int64_t *pQuizIds) override;
  PqaError *ClearOldQuizzes(int64_t maxCount, double maxAgeSec) override;

  // Maintenance mode (BaseEngine.cpp:640-779): the KB is gathered from the shards into ONE un-sharded engine on the first
  // shard's device (device-to-device copies over NVLink), every maintenance call runs there, and FinishMaintenance splits the
  // -- possibly resized -- KB over freshly created shards. Shards hold no gaps, so a KB with removed questions / targets must
  // be compacted before FinishMaintenance.
  PqaError *StartMaintenance(bool forceQuizzes) override;
  PqaError *FinishMaintenance() override;
  PqaError *AddQsTs(int64_t nQuestions, CiAddQorTParam *pAqps, int64_t nTargets, CiAddQorTParam *pAtps) override;
  PqaError *RemoveQuestions(int64_t nQuestions, const int64_t *pQIds) override;
  PqaError *RemoveTargets(int64_t nTargets, const int64_t *pTIds) override;
  PqaError *Compact(int64_t *pnQuestions, const int64_t **ppOldQuestions, int64_t *pnTargets, const int64_t **ppOldTargets) override;

  // single-engine features a group does not offer
  PqaError *SaveKBShard(const char *, bool) override { return No("SaveKBShard (use SaveKB)"); }
  PqaError *SetQuizPriors(int64_t, const double *) override { return No("SetQuizPriors"); }
  PqaError *EvalQuestions(int64_t, const int64_t *, double *, double *, double *, int64_t *) override { return No("EvalQuestions"); }
  PqaError *EvalQuestionsDetailed(int64_t, double *, double *, double *, double *, double *) override { return No("EvalQuestionsDetailed"); }
  PqaError *EvalQuestionsDetailedBatch(int64_t, const int64_t *, double *, double *, double *, double *, double *) override { return No("EvalQuestionsDetailedBatch"); }
  PqaError *ShardEval(int64_t, const int64_t *) override { return No("Shard* protocol (the group drives its shards itself)"); }
  PqaError *ShardSelect(int64_t, const int64_t *, const uint64_t *, int64_t *, void **) override { return No("Shard* protocol"); }
  PqaError *ShardRecordAnswerBegin(int64_t, const int64_t *, const int64_t *) override { return No("Shard* protocol"); }
  PqaError *ShardRecordAnswerEnd(int64_t, const int64_t *) override { return No("Shard* protocol"); }
  PqaError *ShardBuffer(int32_t, void **, int64_t *) override { return No("Shard* protocol"); }
  PqaError *TShardEvalW(int64_t, const int64_t *) override { return No("TShard* protocol"); }
  PqaError *TShardEvalHVL(int64_t, const int64_t *) override { return No("TShard* protocol"); }
  PqaError *TShardPriority(int64_t, const int64_t *) override { return No("TShard* protocol"); }
  PqaError *P2PInit(int32_t, int32_t, int64_t, void **, int64_t *) override { return No("P2P* protocol"); }
  PqaError *P2PExportHandle(uint8_t *) override { return No("P2P* protocol"); }
  PqaError *P2POpenHandle(const uint8_t *, void **) override { return No("P2P* protocol"); }
  PqaError *P2PConnect(void *const *) override { return No("P2P* protocol"); }
  PqaError *P2PNextQuestionBegin(int64_t, const int64_t *, const uint64_t *) override { return No("P2P* protocol"); }
  PqaError *P2PNextQuestionEnd(int64_t, const int64_t *, int64_t *, void **) override { return No("P2P* protocol"); }
  PqaError *P2PRecordAnswerBegin(int64_t, const int64_t *, const int64_t *) override { return No("P2P* protocol"); }
  PqaError *P2PRecordAnswerEnd() override { return No("P2P* protocol"); }
  PqaError *P2PLastPhaseMs(double *) override { return No("P2P* protocol"); }
  PqaError *AnomalyCounts(uint64_t *pCounts3) override;
  PqaError *P2PSetExactOrder(int32_t) override { return No("P2P* protocol (CiB200GroupOptions::_exactOrder)"); }
  PqaError *ResidentBind(int64_t, const int64_t *, const uint64_t *) override { return No("resident stepping"); }
  PqaError *ResidentStep() override { return No("resident stepping"); }
  PqaError *ResidentFetch(int64_t *) override { return No("resident stepping"); }
  double ResidentLastEvalMs() override { return -1.0; }
  PqaError *FlushL2() override { return No("FlushL2"); }

  int shardCount() const { return (int)shards_.size(); }

 private:
  ShardGroup(const CiEngineDefinition &def, const CiB200Options &opts) : Engine(def, opts, ShellTag{}) {}   // shards added by LoadKBGroup
  static PqaError *No(const char *what) { return ErrNotImplemented(std::string("sharded engine group: ") + what); }
  void Connect(const CiB200GroupOptions &gopts);
  void MirrorMaintenanceState();      // dimensions, gap sets and id maps of the maintenance engine -> the shell
  void MoveCells(Engine &whole, bool toWhole);   // shards <-> the un-sharded maintenance engine, device to device
  std::unique_ptr<Engine> maint_;     // exists in maintenance mode only
  CiB200Options baseOpts_;
  CiB200GroupOptions groupOpts_;
  void MirrorResumed(int64_t n, const int64_t *pCounts, const CiAnsweredQuestion *pAQs, const int64_t *pQuizIds);
  static std::vector<CiB200Options> ShardOptions(const CiEngineDefinition &def, const CiB200Options &opts,
                                                 const CiB200GroupOptions &gopts);
  // Everything is validated on the shell before any shard is touched, so an exchanged call can only fail asymmetrically on
  // a device error (a CUDA failure or a barrier time-out on one shard). The shards are then out of lockstep -- epochs, inbox
  // parities, possibly registries -- and the group refuses further exchanged calls instead of answering from diverged state.
  bool broken_ = false;
  static PqaError *Broken() {
    return MakeError(ErrCode::Internal, "sharded engine group: an exchanged call failed on one shard earlier; the shards are out of "
                                        "lockstep. Save the KB if needed and release the engine.");
  }
  std::vector<std::unique_ptr<Engine>> shards_;
  int64_t cap_ = 256;          // quizzes per exchanged call (inbox size); longer batches are cut into slices
};

} // namespace pqa
