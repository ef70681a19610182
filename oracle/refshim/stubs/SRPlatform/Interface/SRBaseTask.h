// TEST INFRASTRUCTURE stub: fork-join task base without the Win32 condition variable
// (replaces /root/reference/ProbQA/SRPlatform/Interface/SRBaseTask.h). The stub thread pool below runs the
// subtasks and joins before EnqueueAdjacent returns, so WaitComplete() has nothing left to wait for.
#pragma once
#include "../SRPlatform/Interface/SRException.h"
#include "../SRPlatform/Interface/SRBasicTypes.h"
namespace SRPlat {
class SRBaseSubtask;
class SRThreadPool;
class SRBaseTask {
public:
  SRBaseTask() {}
  SRBaseTask(const SRBaseTask&) = delete;
  SRBaseTask& operator=(const SRBaseTask&) = delete;
  virtual ~SRBaseTask() {}
  void Reset() {}
  void WaitComplete() {}
  virtual void OnSubtaskComplete(SRBaseSubtask*) {}
  virtual SRThreadPool& GetThreadPool() const = 0;
};
} // namespace SRPlat
