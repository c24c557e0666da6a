#!/bin/bash
# ncu evidence for profiles/ (run on the B200 box through gpurun, ONE GPU):  bash tools/ncu_capture.sh <tag>
#   1. the bench command exits 0 without ncu;
#   2. launch list of the same command (gpu__time_duration per launch, cold-cache and serialised);
#   3. one `--set full` capture of each hot kernel, one read group in read order and 8 read groups segmented.
# Outputs land in gpurun_out/<tag>_*; tools/ncu_summary.py condenses them here afterwards.
set -u
TAG=${1:-r2g}
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-configs --no-fastq --no-cpu --no-e2e"
R8="$B --read-groups 8 --reads 25000000"
$B > $OUT/${TAG}_plain.log 2>&1 || { echo "bench failed without ncu"; exit 1; }
$R8 > $OUT/${TAG}_plain_r8.log 2>&1 || { echo "bench (R = 8) failed without ncu"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_raw.csv $B > $OUT/${TAG}_ncu_l.log 2>&1
for k in build apply; do
  ncu --set full --clock-control none --import-source on -k regex:${k}_smem_kernel -s 3 -c 1 -f -o $OUT/${TAG}_${k}_r1 $B > $OUT/${TAG}_ncu_${k}_r1.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:${k}_smem_kernel -s 3 -c 1 -f -o $OUT/${TAG}_${k}_r8seg $R8 > $OUT/${TAG}_ncu_${k}_r8seg.log 2>&1
done
ls -la $OUT/${TAG}_*
