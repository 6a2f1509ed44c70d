#!/bin/bash
# first GPU session: smoke, parity tests, first bench lines, ncu launch list + one full capture of the top kernel
mkdir -p gpurun_out
exec > >(tee gpurun_out/session1.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv
nproc; free -g | head -2
echo "=== smoke"; time python -c "import __graft_entry__ as g; g.smoke()"
echo "=== pytest gpu"; time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "=== bench default"; timeout 600 python bench.py | tee gpurun_out/bench_fp16s_512.json
for wl in d3q19_srt_fp32_256 d3q19_srt_fp32_512 d3q19_srt_fp16c_512 d3q19_srt_fp16s_256 d3q27_trt_fp32_windtunnel; do
  echo "=== bench $wl"; timeout 600 python bench.py --workload $wl --no-cpu-baseline | tee gpurun_out/bench_$wl.json
done
echo "=== bench general kernel (variant 1)"
for wl in d3q19_srt_fp16s_512 d3q19_srt_fp32_512; do timeout 600 python bench.py --workload $wl --variant 1 --no-cpu-baseline --no-e2e | tee gpurun_out/bench_${wl}_v1.json; done
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fp16s_512.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "=== ncu full, top kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_v4 -s 6 -c 2 -o gpurun_out/prof_fp16s_512 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_collide_v4 -s 6 -c 2 -o gpurun_out/prof_fp32_512 python bench.py --workload d3q19_srt_fp32_512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full32.log 2>&1
ls -la gpurun_out
