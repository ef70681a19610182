// TEST INFRASTRUCTURE stub: swallows log lines, counts them so tests can assert "no warnings fired".
#pragma once
#include "../SRPlatform/Interface/ISRLogger.h"
#include <atomic>
#include <sstream>
namespace SRPlat {
inline std::atomic<long long>& RefShimLogCount() { static std::atomic<long long> n(0); return n; }
class SRLogStream {
  std::ostringstream _os;
public:
  SRLogStream(ISRLogger::Severity, ISRLogger*) {}
  ~SRLogStream() { RefShimLogCount().fetch_add(1); if (getenv("PQA_REF_VERBOSE")) fprintf(stderr, "[ref] %s\n", _os.str().c_str()); }
  template<typename T> SRLogStream& operator<<(const T& v) { _os << v; return *this; }
};
} // namespace SRPlat
