# N-GPU job (N = $NGPU): bench.py under torchrun exactly as the driver launches it, plus the wide-stream worker
set -x
N=${NGPU:-4}
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2>&1 | tail -3
tail -c 800 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/wide_stream_worker.py 2>&1 | tail -2
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().split('\n')[-1])
    print('N',d['n_gpus'],'value',round(d['value']),'S',d['config']['streams_per_gpu'],'B',d['config']['blocks_per_step'],'e2e',round(d['e2e']['value']),'alt',round(d['alt_channelizer']['value']),'numa',d['config'].get('numa_node_of_rank0'))
    for k,v in d['configs'].items():
        if k=='5':
            for r in v['sweep']: print('cfg5 taps',r['taps'],round(r['value']),round(r['ms_per_step'],3),r['ingest_ms_per_step'],r['limited_by'][:12],r['checked'])
        else: print('cfg',k,round(v['value']),round(v['ms_per_step'],3),v.get('ingest_ms_per_step'),v['limited_by'][:12],v['checked'])
except Exception as e: print('parse failed',e)
PY
