#!/usr/bin/env bash
# First GPU run of the experimental row-N1 path: the implicit-conv kernel alone against its numpy
# contract, then the speech-tokenizer decoder against the oracle.  Run on the box:
#   gpurun --timeout 900 -- 'bash tools/n1_bringup.sh'
# Every stage is under its own timeout so a kernel that hangs costs minutes, not the box.
set -u
mkdir -p gpurun_out
export B2A_EXPERIMENTAL_N1=1
{
  echo "== implicit conv kernel"; timeout 300 python -m pytest tests/test_gpu_implicit_conv.py -x -q -m gpu 2>&1 | tail -25
  echo "== qwen3 sampler kernel"; timeout 300 python -m pytest tests/test_gpu_qwen3_sampler.py -x -q -m gpu 2>&1 | tail -25
  echo "== speech tokenizer decoder (bf16 hi/lo operands)"; timeout 500 python -m pytest tests/test_gpu_qwen3_tts_codec.py -x -q -m gpu 2>&1 | tail -40
  echo "== speech tokenizer decoder (fp16 hi/lo operands)"; B2A_ST_FP16=1 timeout 500 python -m pytest tests/test_gpu_qwen3_tts_codec.py -x -q -m gpu 2>&1 | tail -40
} | tee gpurun_out/n1_bringup.log
