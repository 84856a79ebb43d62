mkdir -p gpurun_out
(nvidia-smi topo -m; nproc; free -g | head -2) > gpurun_out/r02ab_topo8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ab_bench_8gpu.log 2> gpurun_out/r02ab_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --workload 2.5Gbp_conifer_k25_16GiB_m0 > gpurun_out/r02ab_bench_8gpu_c4.log 2> gpurun_out/r02ab_bench_8gpu_c4.err
python - <<'PY'
import json
for f in ('gpurun_out/r02ab_bench_8gpu.log','gpurun_out/r02ab_bench_8gpu_c4.log'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(d['n_gpus'], d['config']['workload'], d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['e2e'].get('breakdown_ms'), d['breakdown_ms'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r02ab_bench_8gpu.err
