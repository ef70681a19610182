import sys, time, numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa, synth
Q,K,T=1000,5,1000
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K,Q,T,init_amount=0.1), emulated_workers=16, rng_seed=3)
eng.upload_kb(*synth.binary_search_kb(Q,K,T,0.1,3))
q = eng.start_quiz()
for _ in range(300): eng.next_question(q)
