// probqa_b200: error objects crossing the C ABI. Codes, code texts and the "[code] message=[...] [params]" rendering
// follow the reference (PqaCore/Interface/PqaErrors.h:12-38, PqaCore/PqaErrors.cpp:13-70,127-143,
// PqaCore/PqaErrorParams.cpp). Errors are allocated and released inside this library (CiReleasePqaError).
#pragma once
#include <stdint.h>
#include <string>

namespace pqa {

enum class ErrCode : int64_t {
  None = 0, NotImplemented = 1, SRException = 2, StdException = 3, InsufficientEngineDimensions = 4,
  MaintenanceModeChangeInProgress = 5, MaintenanceModeAlreadyThis = 6, ObjectShutDown = 7, IndexOutOfRange = 8,
  Internal = 9, Aggregate = 10, NegativeCount = 11, NonPositiveAmount = 12, AbsentId = 13, WrongMode = 14,
  UnhandledCase = 15, I64Underflow = 16, QuestionsExhausted = 17, NoQuizActiveQuestion = 18, CantOpenFile = 19,
  FileOp = 20, QuizzesActive = 21, NullArgument = 22, WrongRuntimeType = 23, NotInitialized = 24
};

const char *ErrCodeText(ErrCode c);

struct PqaError {
  ErrCode code;
  std::string message;
  std::string params;   // rendered IPqaErrorParams::ToString(), empty = nullptr params
  bool hasParams;
  std::string ToString(bool withParams) const;
};

PqaError *MakeError(ErrCode code, const std::string &message);
PqaError *MakeError(ErrCode code, const std::string &message, const std::string &params);
// the parameter renderings of PqaErrorParams.cpp
PqaError *ErrIndexOutOfRange(int64_t subject, int64_t first, int64_t last, const std::string &message);
PqaError *ErrAbsentId(int64_t id, const std::string &message);
PqaError *ErrNegativeCount(int64_t count, const std::string &message);
PqaError *ErrNonPositiveAmount(double amount, const std::string &message);
PqaError *ErrNoQuizActiveQuestion(int64_t iAnswer, const std::string &message);
PqaError *ErrNotImplemented(const std::string &feature);
PqaError *ErrInsufficientDims(int64_t nAnswers, int64_t nQuestions, int64_t nTargets);
PqaError *ErrCuda(int cudaError, const char *what, const char *file, int line);
PqaError *ErrStd(const std::string &what);

} // namespace pqa
