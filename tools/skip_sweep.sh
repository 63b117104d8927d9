#!/bin/bash
# decode-step time with parts of the step skipped (timing ablation only; results invalid) -- B2A_SKIP, llama.cu run_layers
for m in none "$@"; do
  echo -n "SKIP=$m "; B2A_SKIP=$m python tools/profile_step.py 320 60 2>&1 | tail -1
done
