# Round-end validation on the GPU box: full -m gpu suite, smoke(), both bench arms with default flags, fresh launch lists.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/final_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/final_smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_step_launches.csv python tools/profile_step.py 320 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/final_whisper_launches.csv python tools/profile_whisper.py 16 8 > /dev/null 2>&1
cat gpurun_out/final_pytest.log gpurun_out/final_smoke.log
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('b200', d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['cpu_baseline']['value'], {k: (d[k]['value'], d[k]['ms_per_step']) for k in ('whisper','snac','qwen3') if k in d})
d=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['cores'], d['ms_per_step'])
"
