// probqa_b200: error objects (see pqa_errors.h). Renderings follow PqaCore/PqaErrors.cpp:13-70,127-143 and the
// ToString() bodies of PqaCore/Interface/PqaErrorParams.h.
#include "pqa_errors.h"

#include <cuda_runtime.h>
#include <sstream>

namespace pqa {

const char *ErrCodeText(ErrCode c) {
  switch (c) {
    case ErrCode::None: return "Success";
    case ErrCode::NotImplemented: return "Not implemented";
    case ErrCode::SRException: return "SRException";
    case ErrCode::StdException: return "std::exception";
    case ErrCode::InsufficientEngineDimensions: return "Insufficient engine dimensions";
    case ErrCode::MaintenanceModeChangeInProgress: return "Maintenance mode change is in progress";
    case ErrCode::MaintenanceModeAlreadyThis: return "Maintenance mode is already this";
    case ErrCode::ObjectShutDown: return "Object is shut(ting) down";
    case ErrCode::IndexOutOfRange: return "Index is out of range";
    case ErrCode::Internal: return "Internal error";
    case ErrCode::Aggregate: return "Aggregate error";
    case ErrCode::NegativeCount: return "The count is negative";
    case ErrCode::NonPositiveAmount: return "The amount is not positive";
    case ErrCode::AbsentId: return "The ID is absent from KB";
    case ErrCode::WrongMode: return "An attempt to execute an operation in a wrong mode";
    case ErrCode::UnhandledCase: return "Unhandled case";
    case ErrCode::I64Underflow: return "Underflow of a 64-bit integer";
    case ErrCode::QuestionsExhausted: return "Engine has run out of questions";
    case ErrCode::NoQuizActiveQuestion: return "No active question in the quiz";
    case ErrCode::CantOpenFile: return "Cannot open file";
    case ErrCode::FileOp: return "File operation failed";
    case ErrCode::QuizzesActive: return "There are still active quizzes";
    case ErrCode::NullArgument: return "Expected non-null argument";
    case ErrCode::WrongRuntimeType: return "Wrong runtime type";
    case ErrCode::NotInitialized: return "Not initialized";
  }
  return "Unknown error code";
}

std::string PqaError::ToString(bool withParams) const {
  std::string s = "[";
  s += ErrCodeText(code);
  s += "] message=[";
  s += message;
  if (!withParams) return s + "]";
  s += "] [";
  s += hasParams ? params : std::string("nullptr");
  s += "]";
  return s;
}

PqaError *MakeError(ErrCode code, const std::string &message) {
  return new PqaError{code, message, std::string(), false};
}
PqaError *MakeError(ErrCode code, const std::string &message, const std::string &params) {
  return new PqaError{code, message, params, true};
}

template <typename T> static std::string str(const T &v) { std::ostringstream o; o.precision(17); o << v; return o.str(); }

PqaError *ErrIndexOutOfRange(int64_t subject, int64_t first, int64_t last, const std::string &message) {
  return MakeError(ErrCode::IndexOutOfRange, message, "subjIndex=" + str(subject) + " not in " + str(first) + "..." + str(last));
}
PqaError *ErrAbsentId(int64_t id, const std::string &message) {
  return MakeError(ErrCode::AbsentId, message, "id=" + str(id));
}
PqaError *ErrNegativeCount(int64_t count, const std::string &message) {
  return MakeError(ErrCode::NegativeCount, message, "count=" + str(count));
}
PqaError *ErrNonPositiveAmount(double amount, const std::string &message) {
  return MakeError(ErrCode::NonPositiveAmount, message, "amount=" + str(amount));
}
PqaError *ErrNoQuizActiveQuestion(int64_t iAnswer, const std::string &message) {
  return MakeError(ErrCode::NoQuizActiveQuestion, message, "answerId=" + str(iAnswer));
}
PqaError *ErrNotImplemented(const std::string &feature) {
  return MakeError(ErrCode::NotImplemented, feature, "Feature=" + feature);
}
PqaError *ErrInsufficientDims(int64_t nAnswers, int64_t nQuestions, int64_t nTargets) {
  // minimums: PqaEngineBaseFactory.h:15-17
  return MakeError(ErrCode::InsufficientEngineDimensions, "",
                   "[nAnswers=" + str(nAnswers) + " of 2] [nQuestions=" + str(nQuestions) + " of 1] [nTargets=" +
                       str(nTargets) + " of 2]");
}
PqaError *ErrCuda(int cudaError, const char *what, const char *file, int line) {
  std::string m = std::string(file) + "(" + str(line) + "): " + what + " failed: " +
                  cudaGetErrorName((cudaError_t)cudaError) + " (" + cudaGetErrorString((cudaError_t)cudaError) + ")";
  return MakeError(ErrCode::SRException, m, "ExceptionType=CudaException");
}
PqaError *ErrStd(const std::string &what) {
  return MakeError(ErrCode::StdException, what, "ExceptionType=std::exception");
}

} // namespace pqa
