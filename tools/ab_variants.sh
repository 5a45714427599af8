#!/bin/bash
# Same-box A/B of experiment builds (tools/build_variant.py): every variants/*.so is swapped in as libb200fe.so and
# timed with tools/time_step.py, REPS times, round-robin so that drift hits all variants alike.
#   bash tools/ab_variants.sh [reps] [variant names...]
reps=${1:-2}; shift
names=("$@"); if [ ${#names[@]} -eq 0 ]; then for f in variants/*.so; do names+=("$(basename $f .so)"); done; fi
cp toolbox_for_asr_and_tts_b200/libb200fe.so /tmp/libb200fe_keep.so
for r in $(seq $reps); do
  for n in "${names[@]}"; do
    cp variants/$n.so toolbox_for_asr_and_tts_b200/libb200fe.so
    echo -n "$n: "; timeout 120 python tools/time_step.py 200 2>&1 | tail -1
  done
done
cp /tmp/libb200fe_keep.so toolbox_for_asr_and_tts_b200/libb200fe.so
