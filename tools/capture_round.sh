#!/bin/bash
# tools/capture_round.sh TAG : the measurement set of a round (run under gpurun, 1 GPU): smoke, bench lines of every
# BASELINE config that fits one GPU, the reference arm, the ncu launch list of the bench command and one
# `ncu --set full` capture of the three marching kernels.  Outputs under gpurun_out/ (copy what is to be judged into profiles/).
TAG="$1"; O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > $O/bench_${TAG}_3d512.json 2> $O/bench_${TAG}.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2>> $O/bench_${TAG}.err
python bench.py --workload 3d512 --profile compact --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_3d512_compact.json 2>> $O/bench_${TAG}.err
python bench.py --workload 2d4096 --steps 20 --warmup 3 > $O/bench_${TAG}_2d4096.json 2>> $O/bench_${TAG}.err
python bench.py --workload m3d512 --steps 5 --warmup 3 > $O/bench_${TAG}_m3d512.json 2>> $O/bench_${TAG}.err
python bench.py --workload sph1536 --profile compact --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_sph1536_compact.json 2>> $O/bench_${TAG}.err
python tools/time_legacy.py 256 > $O/legacy_${TAG}_256.json 2>> $O/bench_${TAG}.err      # legacy MEG6: fused vs pass-by-pass
python tools/time_generic.py 128 > $O/generic_${TAG}_128.json 2>> $O/bench_${TAG}.err   # NVRTC generic mapping vs tables
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_ncu_launches_3d512.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_face --launch-count 3 -f -o $O/prof_${TAG}_3d512 \
  python tools/time_refresh.py 512 fused > $O/ncu_full_$TAG.log 2>&1
ls -la $O/prof_${TAG}_3d512.ncu-rep
for f in 3d512 reference 3d512_compact 2d4096 m3d512 sph1536_compact; do tail -1 $O/bench_${TAG}_$f.json | cut -c1-160; done
