#!/bin/bash
# One GPU-box session of a round: GPU tests, the bench (both arms), the ncu launch list of the bench command and a
# full-set capture of every kernel at the bench size.  Everything lands in gpurun_out/ under the given tag.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh r02
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests_$TAG.log 2>&1; tail -3 $O/gpu_tests_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2> $O/bench_${TAG}_reference_arm.err; echo "reference arm rc=$?"
python bench.py --steps 10 --warmup 3 > $O/bench_${TAG}_N1.json 2> $O/bench_${TAG}_N1.err; echo "bench rc=$?"; tail -4 $O/bench_${TAG}_N1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --configs "" --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"k_enc_|k_dec_|k_scan" -c 45 -o /tmp/ncu_full_$TAG \
  python tools/profile_run.py 3600 2 > $O/ncu_full_$TAG.log 2>&1; echo "full set rc=$?"
ncu -i /tmp/ncu_full_$TAG.ncu-rep --page raw --csv > $O/raw_$TAG.csv 2>/dev/null
for k in k_enc_ltcorr_mma k_enc_pack_rice k_enc_ltlms k_enc_ricetrace k_dec_block k_enc_lagsums k_enc_parcor k_enc_ltfft; do
  ncu -i /tmp/ncu_full_$TAG.ncu-rep --page source --csv -k regex:$k > $O/src_${TAG}_$k.csv 2>/dev/null
done
ls -la $O | tail -5
