cat > /tmp/t2.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa
for (Q,K,T,B) in ((1000,5,1000,256),(2000,5,100000,64)):
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K,Q,T,init_amount=0.1), emulated_workers=16, rng_seed=3, initial_quiz_capacity=B)
    eng.fill_binary_search_kb(3)
    ids = eng.start_quiz_batch(B)
    qs = np.arange(B) % Q
    for _ in range(3):
        eng.set_active_question_batch(ids, qs); eng.record_answer_batch(ids, qs % K); eng.list_top_targets_batch(ids, 10)
        qs = (qs + 7) % Q
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_update|k_normalise|k_list_top|k_set_active" --csv python /tmp/t2.py 2>&1 | grep '"gpu__time' | awk -F'","' '{print $5, $(NF)}' | sed 's/(pqa::DeviceKB.*)//' 
