"""CPU-only checks of the drop-in boundary: libPqaCore.so loads without a GPU, exports every symbol declared in
include/PqaCInterop.h and include/PqaB200Ext.h, its structs have the reference's packing
(PqaCore/Interface/PqaCInterop.h:9-42), and creating an engine without a CUDA device fails loudly with an error
object (no CPU fallback). No compute entry point is called here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pqa():
    from probqa_b200 import build, engine
    build.build()
    engine.load_library()
    return engine


def declared_symbols():
    names = []
    for h in ("PqaCInterop.h", "PqaB200Ext.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"#define\s+PQACORE_API.*", "", src)
        names += re.findall(r"PQACORE_API\s+[^;{]*?\b(\w+)\s*\(", src)
    return names


def test_every_declared_symbol_is_exported_and_bound(pqa):
    names = declared_symbols()
    assert len(names) >= 70, names
    lib = C.CDLL(pqa.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    unbound = [n for n in names if n not in pqa.SIGNATURES]
    assert not unbound, unbound
    # the 40 entry points of the reference header are all there (PqaCInterop.h:48-108)
    ref_names = [n for n in names if not n.startswith("PqaB200_")]
    for n in ("CiGetPqaEngineFactory", "PqaEngineFactory_CreateCpuEngine", "PqaEngine_StartQuiz", "PqaEngine_NextQuestion",
              "PqaEngine_RecordAnswer", "PqaEngine_ListTopTargets", "PqaEngine_RecordQuizTarget", "PqaEngine_Train",
              "PqaEngine_ReleaseQuiz", "CiReleasePqaEngine", "CiReleasePqaError", "PqaError_ToString", "CiReleaseString"):
        assert n in ref_names


def test_struct_layouts_match_reference_packing(pqa):
    # #pragma pack(push, 8): CiEngineDefinition = 3*i64, u8, u16 (offset 26), u32 (28), f64 (32), u64 (40) = 48 bytes
    d = pqa.CiEngineDefinition
    assert C.sizeof(d) == 48
    assert (d.precType.offset, d.precExponent.offset, d.precMantissa.offset, d.initAmount.offset, d.memPoolMaxBytes.offset) == (24, 26, 28, 32, 40)
    assert C.sizeof(pqa.CiAnsweredQuestion) == 16 and C.sizeof(pqa.CiEngineDimensions) == 24
    assert C.sizeof(pqa.CiRatedTarget) == 16 and pqa.CiRatedTarget.prob.offset == 8
    assert pqa.RATED_DTYPE.itemsize == 16


def test_error_objects_roundtrip_without_gpu(pqa):
    lib = pqa.load_library()
    fac = lib.CiGetPqaEngineFactory()
    assert fac
    # insufficient dimensions are rejected before any device work (PqaEngineBaseFactory.cpp:113-122)
    e = C.c_void_p()
    bad = pqa.EngineDefinition(1, 0, 1).to_c()
    assert not lib.PqaEngineFactory_CreateCpuEngine(fac, C.byref(e), C.byref(bad))
    err = pqa.PqaError(e.value)
    assert err.to_string(True) == ("[Insufficient engine dimensions] message=[] "
                                   "[[nAnswers=1 of 2] [nQuestions=0 of 1] [nTargets=1 of 2]]")
    assert err.to_string(False) == "[Insufficient engine dimensions] message=[]"
    # float precision is not accepted, like the reference's CPU factory (PqaEngineBaseFactory.cpp:16-27)
    e = C.c_void_p()
    flt = pqa.EngineDefinition(5, 10, 10, prec_type=pqa.PrecisionType.FLOAT).to_c()
    assert not lib.PqaEngineFactory_CreateCpuEngine(fac, C.byref(e), C.byref(flt))
    assert "[Not implemented]" in pqa.PqaError(e.value).to_string(True)
    # NULL engine handles produce error objects, not crashes
    err = pqa.PqaError(lib.PqaEngine_RecordAnswer(None, 0, 0))
    assert "[Expected non-null argument]" in err.to_string(False)


def test_no_cpu_fallback(pqa):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pqa.PqaException) as ei:
        pqa.PqaEngineFactory().create_cpu_engine(pqa.EngineDefinition(5, 10, 10))
    assert "SRException" in str(ei.value) or "std::exception" in str(ei.value)


def test_product_does_not_touch_the_oracle():
    """The product package must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "probqa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                code = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("#", "//", "*", '"""')))
                assert "import oracle" not in code and "from oracle" not in code and "libpqa_oracle" not in code \
                    and "libpqa_ref" not in code, os.path.join(dirpath, f)


def test_host_logic_self_test_without_gpu(pqa):
    """Gap sets (GapTracker.h), permanent ids (PermanentIdManager.cpp) and the compaction plans of CpuEngine::CompactSpec
    (CpuEngine.cpp:594-641) checked inside the library against brute-force expectations -- host code, no device."""
    import ctypes as C
    lib = pqa.load_library()
    msg = lib.PqaB200_HostLogicSelfTest()
    if msg:
        text = C.string_at(msg).decode()
        lib.CiReleaseString(C.c_void_p(msg))
        raise AssertionError(text)


def test_headers_are_plain_c_and_a_c_client_links(tmp_path):
    """include/*.h must be consumable by a C compiler (the drop-in boundary is a C ABI: no C++ or torch types in the
    signatures), and a C translation unit using both headers must link against libPqaCore.so."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "client.c"
    src.write_text(
        '#include "PqaCInterop.h"\n#include "PqaB200Ext.h"\n#include <stddef.h>\n'
        'int main(void) {\n'
        '  CiEngineDefinition def = {0}; CiB200Options o = {0}; CiB200GroupOptions g = {0};\n'
        '  def._nAnswers = 5; o._device = -1; g._nShards = 2;\n'
        '  /* never called without a GPU: only has to compile and link */\n'
        '  if (def._nAnswers == 0) { void *err = NULL; void *e = PqaB200_CreateShardedEngine(&err, &def, &o, &g);\n'
        '    PqaEngine_StartQuiz(e, &err); PqaEngine_NextQuestionBatch(e, 0, NULL, NULL, NULL, NULL); CiReleasePqaEngine(e); }\n'
        '  return PqaB200_BuildInfo() == NULL;\n}\n')
    from probqa_b200 import build
    lib_dir = os.path.dirname(build.build())
    exe = tmp_path / "client"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src),
                           "-o", str(exe), "-L" + lib_dir, "-lPqaCore", "-Wl,-rpath," + lib_dir])
    assert subprocess.run([str(exe)]).returncode == 0
