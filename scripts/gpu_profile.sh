#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of bench.py's timed region + one --set full capture of the per-step and rebuild kernels.
# Only small CSV exports are kept in gpurun_out/ (the merge-back limit is 64 MiB); usage: scripts/gpu_profile.sh <tag>
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu --no-configs --e2e-steps 3 --ncu-range > $OUT/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_contact_ss|k_reduce|k_leapfrog|k_contact_ws" -c 8 -f -o /tmp/ncu/step python scripts/profile_dense.py 1000000 3000 2 > $OUT/${TAG}_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_pairs_ss|k_carry|k_row_sort|k_csr_fill|k_reorder|k_was_real|k_orphans|k_cellkeys|k_realbits" -c 12 -f -o /tmp/ncu/rebuild python scripts/profile_dense.py 1000000 3000 20 > $OUT/${TAG}_ncu_rebuild.log 2>&1
for r in step rebuild; do
  ncu -i /tmp/ncu/$r.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$r.ncu-rep --page source --csv --print-source cuda,sass > $OUT/${TAG}_${r}_source.csv 2>/dev/null
  sz=$(stat -c %s /tmp/ncu/$r.ncu-rep); echo "$r.ncu-rep $sz bytes"
  if [ "$sz" -lt 20000000 ]; then cp /tmp/ncu/$r.ncu-rep $OUT/${TAG}_$r.ncu-rep; fi
done
gzip -f $OUT/${TAG}_*_source.csv
ls -la $OUT | tail -12
