// TEST INFRASTRUCTURE stub: the RDRAND-seeded xorshift generator is not on the parity path (the random draw is
// injected by the tests); only the declaration that SRDoubleNumber::MakeRandom names is needed.
#pragma once
namespace SRPlat {
class SRFastRandom {
public:
  template<typename T> T Generate();
};
} // namespace SRPlat
