// TEST INFRASTRUCTURE stub of BaseCpuEngine/BaseEngine: just the state the hot-path subtasks read.
#pragma once
#include "../PqaCore/Interface/PqaCommon.h"
#include "../PqaCore/GapTracker.h"
namespace ProbQA {
class BaseCpuEngine {
public:
  EngineDimensions _dims;
  GapTracker<TPqaId> _questionGaps, _targetGaps;
  SRPlat::SRThreadPool _tpWorkers;
  SRPlat::SRBaseMemPool _memPool;
  SRPlat::SRThreadCount _nLooseWorkers;
  explicit BaseCpuEngine(const EngineDimensions& dims, const SRPlat::SRThreadCount nWorkers)
    : _dims(dims), _tpWorkers(nWorkers), _nLooseWorkers(std::max<SRPlat::SRThreadCount>(1, nWorkers - 1)) {}
  virtual ~BaseCpuEngine() {}
  const EngineDimensions& GetDims() const { return _dims; }
  SRPlat::SRThreadPool& GetWorkers() { return _tpWorkers; }
  SRPlat::SRBaseMemPool& GetMemPool() { return _memPool; }
  SRPlat::SRThreadCount GetNLooseWorkers() const { return _nLooseWorkers; }
  const GapTracker<TPqaId>& GetQuestionGaps() const { return _questionGaps; }
  const GapTracker<TPqaId>& GetTargetGaps() const { return _targetGaps; }
  SRPlat::ISRLogger* GetLogger() const { return nullptr; }
};
} // namespace ProbQA
