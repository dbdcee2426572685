#!/bin/bash
# GPU box: one `ncu --set full` capture of the first (= root level, whole batch) launch of every hot
# kernel besides the scan, on the bench workload and on the deep locus; summaries via ncu_summary.py.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
for k in kmer_kernel kmer_fill_kernel kmeans_kernel refcheck_kernel dedupe_kernel partition_kernel unpack_kernel demote_kernel members_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^$k -c 1 \
      -o gpurun_out/ncu_${k}_$TAG -f python scripts/build_once.py bench > gpurun_out/ncu_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/ncu_${k}_$TAG.ncu-rep > gpurun_out/ncu_${k}_$TAG.txt 2>&1
done
for k in kmer_big_insert_kernel kmer_big_fill_kernel refcheck_big_majority_kernel refcheck_big_hamming_kernel dedupe_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^$k -c 1 \
      -o gpurun_out/ncu_deep_${k}_$TAG -f python scripts/build_once.py deep > gpurun_out/ncu_deep_$k.log 2>&1
  python scripts/ncu_summary.py gpurun_out/ncu_deep_${k}_$TAG.ncu-rep > gpurun_out/ncu_deep_${k}_$TAG.txt 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_w1_$TAG.csv python scripts/build_once.py bench > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_deep_$TAG.csv python scripts/build_once.py deep > /dev/null 2>&1
ls -la gpurun_out | tail -30
