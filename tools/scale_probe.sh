#!/bin/bash
# bench.py at N GPUs: the default run (all legs) + the exchange / schedule variants of the sharded symmetric build
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
show() { python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']
        print('$1', 'value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'frac %.3f' % r['frac'], 'kernel_ms %.2f (max %.2f)' % (r['kernel_ms'], r['kernel_ms_max_over_ranks']), 'share %.3f' % r['kernel_share_of_step'], 'launches', r['gemm_launches_per_step'], d['exchange'], 'parity', d['parity_checked'], 'e2e', d['e2e'] and '%.3g' % d['e2e']['value'])
        for k,v in (d.get('kernels') or {}).items():
            if k!='measured_issue_peaks': print('   ', k, '%.4g entries/s' % v['entries_per_s'], 'ms %.2f' % v['ms'], 'frac', v.get('frac'), 'parity', v.get('parity_checked'))
"; }
timeout 600 $T --steps 10 2> gpurun_out/scale_n$N.err | tee gpurun_out/scale_n$N.json | show default
KMG_SHARD_SPLIT=1 timeout 300 $T --steps 10 --no-extras --no-e2e 2>/dev/null | show split1
KMG_SHARD_SPLIT=4 timeout 300 $T --steps 10 --no-extras --no-e2e 2>/dev/null | show split4
timeout 300 $T --steps 10 --no-extras --no-e2e --exchange direct 2>/dev/null | show direct
timeout 300 $T --steps 10 --no-extras --no-e2e --no-sym 2>/dev/null | show plain
