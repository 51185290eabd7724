set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gputests.log
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool python tools/sanitize_workload.py > gpurun_out/san_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/san_$tool.log
done
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 2 -c 1 -o gpurun_out/fe_prof tools/harness 64 128 6 > gpurun_out/fe_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:augment_staged -s 2 -c 1 -o gpurun_out/aug_prof tools/aug_harness 64 0 6 > gpurun_out/aug_ncu.log 2>&1
tools/aug_harness 64 0 20; tools/aug_harness 256 0 10
tail -3 gpurun_out/gputests.log; for t in memcheck racecheck synccheck; do tail -4 gpurun_out/san_$t.log; done
