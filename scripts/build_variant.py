"""Experiment builds: python scripts/build_variant.py NAME -DPQA_EXP_X=1 ...  ->  probqa_b200/lib/exp/NAME/libPqaCore.so
(pqa_eval_staged.cu recompiled with the extra flags, linked with the objects of the regular build). Select at run time
with PQA_B200_LIB=<that path>. Experiments only: nothing in the product refers to these libraries."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probqa_b200 import build as b

name, flags = sys.argv[1], sys.argv[2:]
b.build()
out = os.path.join(b.LIB_DIR, "exp", name)
os.makedirs(out, exist_ok=True)
obj = os.path.join(out, "pqa_eval_staged.o")
subprocess.check_call([b._nvcc()] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + flags +
                      ["-x", "cu", "-c", os.path.join(b.CSRC, "pqa_eval_staged.cu"), "-o", obj])
objs = [os.path.join(b.LIB_DIR, "obj", os.path.splitext(s)[0] + ".o") for s in b.SOURCES if s != "pqa_eval_staged.cu"] + [obj]
lib = os.path.join(out, "libPqaCore.so")
subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs + ["-Xcompiler", "-fPIC"])
print(lib)
