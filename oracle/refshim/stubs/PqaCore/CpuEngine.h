// TEST INFRASTRUCTURE stub of CpuEngine<taNumber>: owns sA/mD/vB as 32-byte aligned, 32-byte padded rows
// (reference layout: CpuEngine.decl.h:31-37 + SRFastArray) and exposes the accessors of CpuEngine.decl.h:77-85.
#pragma once
#include "../PqaCore/BaseCpuEngine.h"
namespace ProbQA {
template<typename taNumber> class CpuEngine : public BaseCpuEngine {
public:
  size_t _ld; // row stride in numbers (multiple of 4)
  taNumber *_sA, *_mD, *_vB;
  CpuEngine(const EngineDimensions& dims, const SRPlat::SRThreadCount nWorkers) : BaseCpuEngine(dims, nWorkers) {
    _ld = (size_t(dims._nTargets) + 3) & ~size_t(3);
    _sA = static_cast<taNumber*>(_mm_malloc(sizeof(taNumber) * _ld * size_t(dims._nQuestions) * size_t(dims._nAnswers), 32));
    _mD = static_cast<taNumber*>(_mm_malloc(sizeof(taNumber) * _ld * size_t(dims._nQuestions), 32));
    _vB = static_cast<taNumber*>(_mm_malloc(sizeof(taNumber) * _ld, 32));
  }
  ~CpuEngine() { _mm_free(_sA); _mm_free(_mD); _mm_free(_vB); }
  const taNumber& GetA(const TPqaId i, const TPqaId k, const TPqaId j) const { return _sA[(size_t(i) * size_t(_dims._nAnswers) + size_t(k)) * _ld + size_t(j)]; }
  taNumber& ModA(const TPqaId i, const TPqaId k, const TPqaId j) { return _sA[(size_t(i) * size_t(_dims._nAnswers) + size_t(k)) * _ld + size_t(j)]; }
  const taNumber& GetD(const TPqaId i, const TPqaId j) const { return _mD[size_t(i) * _ld + size_t(j)]; }
  taNumber& ModD(const TPqaId i, const TPqaId j) { return _mD[size_t(i) * _ld + size_t(j)]; }
  const taNumber& GetB(const TPqaId j) const { return _vB[size_t(j)]; }
  taNumber& ModB(const TPqaId j) { return _vB[size_t(j)]; }
};
} // namespace ProbQA
