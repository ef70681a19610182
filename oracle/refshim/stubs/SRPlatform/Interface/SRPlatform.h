// TEST INFRASTRUCTURE stub: replaces the MSVC-only SRPlatform.h (which #errors on other compilers,
// /root/reference/ProbQA/SRPlatform/Interface/SRPlatform.h:44). Declares nothing numeric.
#pragma once
#define IS_CPU_X86_32 0
#define IS_CPU_X86_64 1
