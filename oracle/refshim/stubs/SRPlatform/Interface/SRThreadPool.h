// TEST INFRASTRUCTURE stub: std::thread fork-join replacement for the Win32 SRThreadPool
// (/root/reference/ProbQA/SRPlatform/SRThreadPool.cpp:59 uses CreateThread). Subtasks are independent, so the
// execution order cannot change any result; only GetWorkerCount() (which sizes the splits) is semantic.
#pragma once
#include "../SRPlatform/Interface/SRBasicTypes.h"
#include "../SRPlatform/Interface/SRBaseTask.h"
#include "../SRPlatform/Interface/SRBaseSubtask.h"
#include <thread>
#include <vector>
#include <atomic>
namespace SRPlat {
class SRThreadPool {
  SRThreadCount _nWorkers;   // what the reference calls hardware_concurrency(): sizes every split
  SRThreadCount _nOsThreads; // how many OS threads actually execute subtasks (1 = inline)
public:
  explicit SRThreadPool(const SRThreadCount nWorkers, const SRThreadCount nOsThreads = 1)
    : _nWorkers(nWorkers), _nOsThreads(nOsThreads) {}
  SRThreadCount GetWorkerCount() const { return _nWorkers; }
  void SetOsThreads(const SRThreadCount n) { _nOsThreads = n; }
  template<typename taSubtask> void EnqueueAdjacent(taSubtask *pFirst, const SRSubtaskCount nSubtasks, SRBaseTask&) {
    if (_nOsThreads <= 1 || nSubtasks <= 1) {
      for (SRSubtaskCount i = 0; i < nSubtasks; i++) pFirst[i].Run();
      return;
    }
    std::atomic<SRSubtaskCount> next(0);
    auto body = [&]() {
      for (;;) {
        const SRSubtaskCount i = next.fetch_add(1);
        if (i >= nSubtasks) return;
        pFirst[i].Run();
      }
    };
    const SRSubtaskCount nThr = std::min<SRSubtaskCount>(_nOsThreads, nSubtasks);
    std::vector<std::thread> th;
    for (SRSubtaskCount t = 1; t < nThr; t++) th.emplace_back(body);
    body();
    for (auto& t : th) t.join();
  }
};
} // namespace SRPlat
