timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err
timeout 200 ncu --set full --clock-control none -k regex:"mha_fwd|mel_log|tc_gemm_kernel" -s 8 -c 14 -o /tmp/wh python tools/profile_whisper.py 16 2 > /dev/null 2>&1
ncu -i /tmp/wh.ncu-rep --page raw --csv > gpurun_out/ncu_whisper.csv 2>/dev/null; rm -f /tmp/wh.ncu-rep
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('b200', d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['cpu_baseline']['value'])
d=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['cores'], d['ms_per_step'])
"
wc -c gpurun_out/ncu_whisper.csv
