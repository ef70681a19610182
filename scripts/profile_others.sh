#!/bin/bash
# ncu captures of the kernels around the evaluation (posterior update, top targets, selection, training) at BASELINE
# config 2 (B = 256) and at T = 100 000 (B = 64); summaries -> gpurun_out/r02_other_kernels.txt
mkdir -p gpurun_out /tmp/prof
cat > /tmp/others.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa, synth
def run(Q, K, T, B):
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=16, rng_seed=3, initial_quiz_capacity=B)
    eng.fill_binary_search_kb(3)
    ids = eng.start_quiz_batch(B)
    r = np.arange(B, dtype=np.uint64) * 7919
    for step in range(3):
        ch = eng.next_question_batch(ids, r)
        eng.record_answer_batch(ids, [int(c) % K for c in ch])
        eng.list_top_targets_batch(ids, 10)
    eng.record_quiz_target_batch(ids, [(7 * x) % T for x in range(B)])
    eng.close()
run(1000, 5, 1000, 256)
run(2000, 5, 100000, 64)
# config 5 shape: one big RecordQuizTarget batch (device-grouped path)
Q, K, T, n, d = 1000, 5, 1000, 20000, 8
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=16, rng_seed=3, initial_quiz_capacity=n)
eng.fill_binary_search_kb(3)
rng = np.random.default_rng(1)
ids = eng.start_quiz_batch(n)
qs = np.argsort(rng.random((n, 64)), axis=1)[:, :d] + rng.integers(0, Q - 64, size=(n, 1))
for s in range(d):
    eng.set_active_question_batch(ids, qs[:, s]); eng.record_answer_batch(ids, rng.integers(0, K, size=n))
eng.record_quiz_target_batch(ids, rng.integers(0, T, size=n))
PY
ncu --set full --clock-control none -k regex:"k_update_priors|k_update_lanes|k_normalise_rows|k_list_top_targets|k_select_question|k_train|k_add_vb|k_build_derived|DeviceRadixSort" -c 80 -o /tmp/prof/others -f python /tmp/others.py > gpurun_out/r02_others_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/prof/others.ncu-rep > gpurun_out/r02_other_kernels_full.txt 2>&1
python - <<'PY'
import re
txt = open('gpurun_out/r02_other_kernels_full.txt').read().split('---- launch id')
keep = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
with open('gpurun_out/r02_other_kernels.txt', 'w') as f:
    for blk in txt[1:]:
        f.write('---- launch' + blk.split('\n')[0] + '\n')
        for ln in blk.split('\n'):
            if any(k in ln for k in keep): f.write(ln + '\n')
        st = blk.split('top warp stall reasons')
        if len(st) > 1: f.write('  top stalls:' + ' '.join(x.strip() for x in st[1].split('\n')[1:4]) + '\n')
PY
wc -l gpurun_out/r02_other_kernels.txt
