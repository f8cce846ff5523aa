# 2-GPU job: multi-device C ABI over two real GPUs, wide-stream worker under torchrun, bench.py at N=2 with all configs
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py::test_wide_stream_channel_sharding -q > gpurun_out/r2_pytest_2gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_2gpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | tail -3
tail -c 1200 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.json').read().strip().split('\n')[-1])
    print('N=2 value',round(d['value']),'S',d['config']['streams_per_gpu'],'e2e',round(d['e2e']['value']),'alt',round(d['alt_channelizer']['value']))
    for k,v in d['configs'].items():
        if k=='5':
            for r in v['sweep']: print('cfg5 taps',r['taps'],round(r['value']),r['ms_per_step'],r['ingest_ms_per_step'],r['limited_by'][:12],r['checked'])
        else: print('cfg',k,round(v['value']),v['ms_per_step'],v.get('ingest_ms_per_step'),v['limited_by'][:12],v['checked'])
except Exception as e: print('parse failed',e)
PY
