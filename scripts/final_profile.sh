mkdir -p gpurun_out
CMD="python bench.py --packets 4e7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1wf8_launches.csv $CMD > gpurun_out/r1wf8_launches_bench.log 2>&1
python bench.py > gpurun_out/r1wf8_bench.json 2> gpurun_out/r1wf8_bench.err
tail -c 300 gpurun_out/r1wf8_bench.json
