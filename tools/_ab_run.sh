timeout 900 python -m pytest tests/test_bench_contract.py -x -q -m gpu > gpurun_out/r02i_pytest_bench.log 2>&1; tail -15 gpurun_out/r02i_pytest_bench.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -5 gpurun_out/r02i_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench.json') if l.startswith('{')][0])
for k in ('value','ms_per_step','kernel_ms','e2e','e2e_full','gpu_launches','strong','single_runs','cpu_baseline','cpu_port','parity'):
    print(k, json.dumps(d[k])[:700])
print('roofline', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['peak'])
print('config5', json.dumps(d['config5'])[:1200])
PY
(time python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/r02i_bench_ref.json 2>> gpurun_out/r02i_bench.err; head -c 1500 gpurun_out/r02i_bench_ref.json; tail -4 gpurun_out/r02i_bench.err
