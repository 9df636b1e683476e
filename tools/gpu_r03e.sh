#!/bin/bash
# round 2, GPU call 3E (8 GPUs): e2e with / without binding each rank to its GPU's NUMA node; full bench line of the final tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r03e_topo.txt
for f in /sys/bus/pci/devices/*/numa_node; do :; done
summ() { python - "$1" <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=b['e2e']['host_phases_ms']['per_rank']
print('value %.4g e2e %.4g ctor ms %s numa %s' % (b['value'], b['e2e']['value'], [round(r['constructor'],1) for r in ph], b['config'].get('host_numa_binding_rank0')))
PY
}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r03e_bind.json 2> gpurun_out/r03e_bind.err; echo "bind rc=$?"; summ gpurun_out/r03e_bind.json
timeout 600 $T --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 --skip-large --skip-cpu --no-numa-bind > gpurun_out/r03e_nobind.json 2> gpurun_out/r03e_nobind.err; echo "nobind rc=$?"; summ gpurun_out/r03e_nobind.json
timeout 600 $T --master-port 29563 bench.py --gpus 8 --steps 20 --warmup 5 --skip-large --skip-cpu > gpurun_out/r03e_bind2.json 2> gpurun_out/r03e_bind2.err; echo "bind2 rc=$?"; summ gpurun_out/r03e_bind2.json
timeout 900 $T --master-port 29564 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r03e_bench8.json 2> gpurun_out/r03e_bench8.err; echo "bench8 rc=$?"; summ gpurun_out/r03e_bench8.json
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r03e_bench8.json').read().strip().splitlines()[-1])
for k in ('batched_strong','sharded_large_n','sharded_check'):
    print(k, json.dumps(b.get(k))[:500])
PY
