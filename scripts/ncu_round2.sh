#!/bin/bash
# GPU box, round 2: `ncu --set full` captures of the kernels DESIGN.md / the bench line cite, on the workloads where
# they matter, plus launch lists; summaries via scripts/ncu_summary.py.  Usage: bash scripts/ncu_round2.sh [tag]
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
cap() {  # cap <kernel regex> <driver script + args> <name>
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 -c 1 \
      -o gpurun_out/ncu_$3_$TAG -f python $2 > gpurun_out/ncu_$3.log 2>&1
  python scripts/ncu_summary.py gpurun_out/ncu_$3_$TAG.ncu-rep > gpurun_out/ncu_$3_$TAG.txt 2>&1
}
# bench workload (BASELINE config #2, 1,000 loci, one device-resident build): root-level launches
cap '^scan_kernel' "scripts/build_once.py bench" scan_kernel
cap '^kmeans_kernel_w32' "scripts/build_once.py bench" kmeans_kernel_w32
cap '^refcheck_kernel' "scripts/build_once.py bench" refcheck_kernel
cap '^dedupe_kernel' "scripts/build_once.py bench" dedupe_kernel
cap '^prg_walk_kernel' "scripts/build_once.py bench" prg_walk_kernel
cap '^expand_partition_kernel' "scripts/build_once.py bench" expand_partition_kernel
cap '^partition_kernel' "scripts/build_once.py bench" partition_kernel
# flat deep locus (10,000 x 20,000): whole-grid one-reference-like check on the packed rows, k-mer numbering
cap '^majority_count_kernel' "scripts/build_flat_once.py" flat_majority_count_kernel
cap '^hamming_packed_kernel' "scripts/build_flat_once.py" flat_hamming_packed_kernel
cap '^kmer_big_insert_kernel' "scripts/build_flat_once.py" flat_kmer_big_insert_kernel
# deep-clade locus (1,500 x 4,000): the CTA-group KMeans
cap '^kmeans_group_kernel' "scripts/build_deep_once.py" deep_kmeans_group_kernel
for w in bench; do
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
      --log-file gpurun_out/launches_bench_$TAG.csv python scripts/build_once.py bench > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/launches_bench_$TAG.csv > gpurun_out/launches_bench_${TAG}_summary.txt
done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_flat_$TAG.csv python scripts/build_flat_once.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/launches_flat_$TAG.csv > gpurun_out/launches_flat_${TAG}_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_deep_$TAG.csv python scripts/build_deep_once.py > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/launches_deep_$TAG.csv > gpurun_out/launches_deep_${TAG}_summary.txt
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out | tail -40
