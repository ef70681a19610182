mkdir -p gpurun_out; OUT=gpurun_out/r02_client.log; : > $OUT
for L in 1 16 64 256; do
  echo "### learners $L" >> $OUT
  PQA_B200_STATS=1 ./probqa_b200/lib/pqa_client --trainings 12000 --learners $L --report-every 4096 --progress /tmp/p$L.txt --kb-dir /tmp >> $OUT 2>&1
done
grep "learners\|questions_per_s\|calls in\|combined" $OUT | cut -c1-260
