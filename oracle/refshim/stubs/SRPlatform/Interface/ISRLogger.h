// TEST INFRASTRUCTURE stub logger interface (severity names from ISRLogger.h:11-26).
#pragma once
namespace SRPlat {
class ISRLogger {
public:
  enum class Severity : unsigned char { Info, Warning, Error, Critical };
  virtual ~ISRLogger() {}
};
} // namespace SRPlat
