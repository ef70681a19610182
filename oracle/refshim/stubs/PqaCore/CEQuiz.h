// TEST INFRASTRUCTURE stub of CEQuiz<taNumber>: the per-quiz arrays of CEQuiz.decl.h:14-45.
#pragma once
#include "../PqaCore/CpuEngine.h"
#include "../PqaCore/CEBaseTask.h"
#include "../PqaCore/CERecordAnswerTask.h"
#include "../PqaCore/CESetPriorsTask.h"
namespace ProbQA {
template<typename taNumber> class CEQuiz {
public:
  typedef int64_t TExponent;
  taNumber *_pPriorMants; TExponent *_pTlhExps; __m256i *_isQAsked;
  size_t _ld, _nAskedVects;
  explicit CEQuiz(const EngineDimensions& dims) {
    _ld = (size_t(dims._nTargets) + 3) & ~size_t(3);
    _nAskedVects = (size_t(dims._nQuestions) + 255) / 256 + 1;
    _pPriorMants = static_cast<taNumber*>(_mm_malloc(sizeof(taNumber) * _ld, 32));
    _pTlhExps = static_cast<TExponent*>(_mm_malloc(sizeof(TExponent) * _ld, 32));
    _isQAsked = static_cast<__m256i*>(_mm_malloc(32 * _nAskedVects, 32));
    memset(_pPriorMants, 0, sizeof(taNumber) * _ld);
    memset(_pTlhExps, 0, sizeof(TExponent) * _ld);
    memset(_isQAsked, 0, 32 * _nAskedVects);
  }
  ~CEQuiz() { _mm_free(_pPriorMants); _mm_free(_pTlhExps); _mm_free(_isQAsked); }
  taNumber* GetPriorMants() const { return _pPriorMants; }
  TExponent* GetTlhExps() const { return _pTlhExps; }
  __m256i* GetQAsked() const { return _isQAsked; }
};
} // namespace ProbQA
