#!/bin/bash
# timing only: tools/debug/critic_time.sh <variant> ...
for v in "$@"; do
  unset OFDMGAN_LIB
  if [ "$v" != default ]; then export OFDMGAN_LIB=$PWD/ofdm-gan-sr_b200/lib/libofdmgan_$v.so; fi
  echo "$v $(python tools/debug/time_train.py 2>&1 | tail -1) $(nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader)"
done
