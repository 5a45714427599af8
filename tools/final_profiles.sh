#!/bin/bash
# One-GPU measurement passes behind the numbers in DESIGN.md section 7; everything lands in gpurun_out/.
#   gpurun --timeout 1700 -- 'bash tools/final_profiles.sh bench'      # tests + all plain bench runs + the launch list
#   gpurun --timeout 600  -- 'bash tools/final_profiles.sh warp|stream|tts'   # one ncu --set full capture each
# A number printed under ncu is never a bench value: every capture follows a plain run of the same command.
set -u
out=gpurun_out
mkdir -p $out
case "${1:-bench}" in
bench)
  (timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $out/final_pytest.txt
  python bench.py --impl reference --steps 3 --warmup 1 > $out/final_bench_reference.json 2> $out/final_bench_reference.err
  python bench.py --steps 20 --warmup 5 > $out/final_bench_n1_steps20.json 2> $out/final_bench_n1.err
  python bench.py --steps 200 --warmup 5 > $out/final_bench_n1_steps200.json 2>> $out/final_bench_n1.err
  python bench_streaming.py > $out/final_bench_streaming_n1.json 2> $out/final_stream.err
  python bench_streaming.py --streams 1024 > $out/final_bench_streaming_n1_1024.json 2>> $out/final_stream.err
  python bench_streaming.py --streams 256 > $out/final_bench_streaming_n1_256.json 2>> $out/final_stream.err
  python bench_tts.py --cpu > $out/final_bench_tts_n1.json 2> $out/final_tts.err
  python bench_bulk.py --hours 125 > $out/final_bench_bulk_n1.json 2> $out/final_bulk.err
  python tools/time_step.py 200 > $out/final_time_step.txt 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/final_launches.csv \
      python bench.py --steps 2 --warmup 1 > $out/final_ncu_launches.log 2>&1
  tail -3 $out/final_pytest.txt
  cut -c1-200 $out/final_bench_n1_steps20.json
  ;;
warp)
  python tools/profile_step.py 6 && ncu --set full --import-source on --clock-control none -k regex:fbank_warp_kernel -s 3 -c 1 \
      -o $out/final_warp python tools/profile_step.py 6 > $out/final_ncu_warp.log 2>&1
  tail -1 $out/final_ncu_warp.log
  ;;
stream)
  python bench_streaming.py --no-graph --ticks 10 --warmup 3 > /dev/null && ncu --set full --import-source on --clock-control none \
      -k regex:stream_push_kernel -s 10 -c 1 -o $out/final_stream python bench_streaming.py --no-graph --ticks 10 --warmup 3 > $out/final_ncu_stream.log 2>&1
  tail -1 $out/final_ncu_stream.log
  ;;
tts)
  python bench_tts.py --steps 2 --warmup 1 > /dev/null && ncu --set full --import-source on --clock-control none -k regex:tts_mel_kernel \
      -s 1 -c 1 -o $out/final_tts python bench_tts.py --steps 2 --warmup 1 > $out/final_ncu_tts.log 2>&1
  tail -1 $out/final_ncu_tts.log
  ;;
esac
