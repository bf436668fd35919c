#!/bin/bash
# Round-2 measurement run on one B200 (under gpurun): parity tests, bench.py on every workload, the reference arm, the ncu
# launch list and full captures of the hot kernels of the same build, compute-sanitizer over small parity cases.
# usage: scripts/final_profile.sh <tag> [parts]      parts: any of tests bench ncu san (default: all)
TAG=${1:-r2final}
PARTS=${2:-"tests bench ncu san"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
if [[ $PARTS == *tests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
  tail -4 gpurun_out/${TAG}_gpu_tests.log
fi
if [[ $PARTS == *bench* ]]; then
  for c in cfg2 cfg1 cfg4 cfg5; do
    timeout 1200 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
    echo "bench $c rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_$c.json
  done
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref_cfg2.json 2> gpurun_out/${TAG}_ref_cfg2.err
  cut -c1-200 gpurun_out/${TAG}_ref_cfg2.json
  # diagnostic, not a BASELINE workload: cfg2 with a second dust component of another material mix (several-component kernels)
  SK_BENCH_SECOND_MIX=1 timeout 600 python bench.py --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_cfg2_second_mix.json 2> gpurun_out/${TAG}_bench_cfg2_second_mix.err
  cut -c1-200 gpurun_out/${TAG}_bench_cfg2_second_mix.json
  # diagnostic: cfg2 with a rotating ring and source (kinematics kernels; path-length stretching is off then, as in the reference)
  SK_BENCH_KINEMATICS=1 timeout 600 python bench.py --no-cpu-baseline --no-parity --packets 3e7 > gpurun_out/${TAG}_bench_cfg2_kinematics.json 2> gpurun_out/${TAG}_bench_cfg2_kinematics.err
  cut -c1-200 gpurun_out/${TAG}_bench_cfg2_kinematics.json
fi
if [[ $PARTS == *ncu* ]]; then
  B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $B --packets 4e7 > gpurun_out/${TAG}_launches_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sk_wf_trace -s 6 -c 2 -o gpurun_out/${TAG}_prof -f $B --packets 2e7 > gpurun_out/${TAG}_prof_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:sk_wf_(launch|advance|detect)" -s 9 -c 3 -o gpurun_out/${TAG}_events -f $B --packets 2e7 > gpurun_out/${TAG}_events_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:sk_wf_(trace|bin_scatter)" -s 6 -c 3 -o gpurun_out/${TAG}_cfg4 -f $B --config cfg4 --packets 2e6 > gpurun_out/${TAG}_cfg4_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sk_wf_trace -s 4 -c 2 -o gpurun_out/${TAG}_cfg1 -f $B --config cfg1 --packets 1e7 > gpurun_out/${TAG}_cfg1_bench.log 2>&1
  SK_BENCH_SITES=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sk_wf_trace -s 4 -c 2 -o gpurun_out/${TAG}_cfg5 -f $B --config cfg5 --packets 4e6 > gpurun_out/${TAG}_cfg5_bench.log 2>&1
  ls -la gpurun_out | grep ${TAG} | grep ncu-rep
  # gpurun brings back at most 64 MiB: the raw metric pages of every capture as CSV (what scripts/make_profile_summaries.py reads),
  # and only the report of the dominant trace kernels itself (source page, scripts/make_traffic.py)
  for r in prof events cfg4 cfg1 cfg5; do
    if [ -f gpurun_out/${TAG}_$r.ncu-rep ]; then
      ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null
      [ $r != prof ] && rm -f gpurun_out/${TAG}_$r.ncu-rep
    fi
  done
  du -sh gpurun_out
fi
if [[ $PARTS == *san* ]]; then
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_setup.py -q -x -k "cartesian_cfg1 or octree_cfg2_small or explicit or interleaved or dust_emission or voronoi or two_components or three_components or particle_density or kinematics" > gpurun_out/${TAG}_memcheck.log 2>&1
  echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.log
  tail -5 gpurun_out/${TAG}_memcheck.log
fi
