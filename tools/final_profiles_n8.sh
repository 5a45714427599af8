#!/bin/bash
# Eight-GPU measurement pass (one box): the driver's launch line for bench.py, then the side benches of configs[2..4],
# each followed by its one-GPU run on the same box (per-GPU efficiency).
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/final_profiles_n8.sh'
set -u
out=gpurun_out
mkdir -p $out
run8() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run8 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $out/n8_bench.json 2> $out/n8_bench.err
run8 29512 bench_streaming.py > $out/n8_streaming.json 2> $out/n8_streaming.err
python bench_streaming.py > $out/n8_streaming_n1_samebox.json 2>> $out/n8_streaming.err
run8 29513 bench_bulk.py --hours 1000 > $out/n8_bulk.json 2> $out/n8_bulk.err
python bench_bulk.py --hours 125 > $out/n8_bulk_n1_samebox.json 2>> $out/n8_bulk.err
run8 29514 bench_tts.py > $out/n8_tts.json 2> $out/n8_tts.err
python bench_tts.py > $out/n8_tts_n1_samebox.json 2>> $out/n8_tts.err
for f in $out/n8_*.json; do echo "$f: $(cut -c1-110 $f | tail -1)"; done
