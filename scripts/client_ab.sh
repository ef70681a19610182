#!/bin/bash
# the reference's client loop with the current library and with probqa_b200/lib/exp/old (same box, alternating)
for rep in 1 2; do for v in new old; do for L in ${LEARNERS:-1 64 256}; do
  if [ $v = old ]; then export LD_LIBRARY_PATH=$PWD/probqa_b200/lib/exp/old; else unset LD_LIBRARY_PATH; fi
  echo "$v learners $L: $(./probqa_b200/lib/pqa_client --trainings 8000 --learners $L --report-every 4096 --progress /tmp/p.txt --kb-dir /tmp 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.read().strip().splitlines()[-1])['questions_per_s'])")"
done; done; done
