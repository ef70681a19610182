#!/bin/bash
# TEST INFRASTRUCTURE: builds oracle/_ref/libpqa_ref.so from the reference's OWN sources under
# /root/reference/ProbQA (hot-path subtask bodies + numeric substrate), for pinning the oracle.
#  * sources are read where they lie; MSVC->GCC token patches (refshim/patch.pl) are applied to a scratch
#    copy under a mktemp dir that is deleted afterwards; nothing of the reference is copied into the repo;
#  * Win32-bound shell headers (thread pool, mem pool, logger, engine/quiz shells) are replaced by the
#    stubs in refshim/stubs/ -- none of them contains arithmetic;
#  * flags follow SURVEY.md 8(c): -O2 -mavx2 -mfma -mbmi -mbmi2 -ffp-contract=off, plus -fno-strict-aliasing: the
#    reference reads vector lanes through MSVC's union members (x.m128i_i64[i]), which the shim turns into pointer casts;
#    MSVC never applies type-based alias analysis, and g++ with it on dropped the store behind such a read in
#    SRSimd::FullHorizMaxI64 (ResumeQuiz then normalised with a wrong common exponent -- invisible in the priors, which
#    are scale-invariant, but visible as -0.0 in removed-target lanes; found by tests/golden, round 1).
# Only the .so lands in oracle/_ref/ (git-ignored, travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PQA_REFERENCE_ROOT:-/root/reference}/ProbQA"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "build_ref: $REF not present; keeping prebuilt $OUT" >&2; exit 0; fi
TMP="$(mktemp -d /tmp/pqa_refbuild.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT" "$TMP/ProbQA/SRPlatform/Interface" "$TMP/ProbQA/PqaCore/Interface" "$TMP/ProbQA/harness"

for f in "$REF"/SRPlatform/Interface/*.h; do "$HERE/refshim/patch.pl" "$f" > "$TMP/ProbQA/SRPlatform/Interface/$(basename "$f")"; done
for f in "$REF"/PqaCore/*.h; do "$HERE/refshim/patch.pl" "$f" > "$TMP/ProbQA/PqaCore/$(basename "$f")"; done
for f in "$REF"/PqaCore/Interface/*.h; do "$HERE/refshim/patch.pl" "$f" > "$TMP/ProbQA/PqaCore/Interface/$(basename "$f")"; done
SR_CPP="SRSimd.cpp SRVectMath.cpp"
PQA_CPP="CEEvalQsSubtaskConsider.cpp CERecordAnswerSubtaskMul.cpp CESetPriorsSubtaskSum.cpp CEHeapifyPriorsSubtaskMake.cpp CEUpdatePriorsSubtaskMul.cpp CENormPriorsSubtaskMax.cpp CENormPriorsSubtaskCorrSum.cpp CEListTopTargetsAlgorithm.cpp CERadixSortRatingsSubtaskSort.cpp CETrainOperation.cpp"
for f in $SR_CPP; do "$HERE/refshim/patch.pl" "$REF/SRPlatform/$f" > "$TMP/ProbQA/SRPlatform/$f"; done
for f in $PQA_CPP; do "$HERE/refshim/patch.pl" "$REF/PqaCore/$f" > "$TMP/ProbQA/PqaCore/$f"; done
# overlay the stubs (they replace the same-named scratch copies)
cp -r "$HERE/refshim/stubs/SRPlatform/." "$TMP/ProbQA/SRPlatform/"
cp -r "$HERE/refshim/stubs/PqaCore/." "$TMP/ProbQA/PqaCore/"
cp "$HERE/ref_harness_prims.cpp" "$HERE/ref_harness_engine.cpp" "$TMP/ProbQA/harness/"
cp "$HERE/refshim/stubs/PqaCore/stdafx.h" "$TMP/ProbQA/harness/stdafx.h"

CXX="${CXX:-g++}"
FLAGS="-std=c++17 -O2 -fPIC -mavx2 -mfma -mbmi -mbmi2 -ffp-contract=off -fno-fast-math -fno-strict-aliasing -fpermissive -w -pthread -include $HERE/refshim/compat.h -I $TMP/ProbQA/PqaCore"
OBJS=""
for f in $SR_CPP; do $CXX $FLAGS -c "$TMP/ProbQA/SRPlatform/$f" -o "$TMP/sr_${f%.cpp}.o"; OBJS="$OBJS $TMP/sr_${f%.cpp}.o"; done
for f in $PQA_CPP; do $CXX $FLAGS -c "$TMP/ProbQA/PqaCore/$f" -o "$TMP/pqa_${f%.cpp}.o"; OBJS="$OBJS $TMP/pqa_${f%.cpp}.o"; done
for f in ref_harness_prims.cpp ref_harness_engine.cpp; do $CXX $FLAGS -c "$TMP/ProbQA/harness/$f" -o "$TMP/h_${f%.cpp}.o"; OBJS="$OBJS $TMP/h_${f%.cpp}.o"; done
$CXX -shared -pthread -o "$OUT/libpqa_ref.so" $OBJS
echo "build_ref: built $OUT/libpqa_ref.so"
