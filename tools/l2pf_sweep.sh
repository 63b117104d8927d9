#!/bin/bash
# decode-step time for several L2-prefetch schedules (B2A_L2PF bit mask, see llama.cu pf_of)
for m in "$@"; do
  echo -n "L2PF=$m "; B2A_L2PF=$m python tools/profile_step.py 320 60 2>&1 | tail -1
done
