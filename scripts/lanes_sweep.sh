#!/bin/bash
# GPU box (2 GPUs): builds in flight per rank (MPRG_BUILD_LANES) x how a lane's host thread waits (MPRG_LANE_WAIT)
# x host cores (taskset), bench.py's value / e2e per step.  Output: profiles/r2_lanes_sweep.txt
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], 'n', d['n_gpus'], 'lanes', d['config']['builds_in_flight'], 'value_ms', round(d['ms_per_step'],3), 'serial_ms', round(d['one_at_a_time']['ms_per_step'],3), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'e2e_serial', round(d['e2e']['one_at_a_time']['ms_per_step'],3))" $1 "$2"; }
run2() { # name, cpu list, env...
  name=$1; cpus=$2; shift; shift
  env "$@" timeout 120 taskset -c $cpus python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 4 --no-files --no-big > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err || tail -c 300 gpurun_out/sw_$name.err
  show gpurun_out/sw_$name.json "$name"
}
run1() { name=$1; cpus=$2; shift; shift
  env "$@" timeout 120 taskset -c $cpus python bench.py --steps 20 --warmup 4 --no-files --no-big > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err || tail -c 300 gpurun_out/sw_$name.err
  show gpurun_out/sw_$name.json "$name"
}
nproc
run2 n2_24c_yield6 0-23 MPRG_LANE_WAIT=yield MPRG_BUILD_LANES=6
run2 n2_8c_spin2 0-7 MPRG_LANE_WAIT=spin MPRG_BUILD_LANES=2
run2 n2_8c_yield3 0-7 MPRG_LANE_WAIT=yield MPRG_BUILD_LANES=3
run1 n1_16c_spin4 0-15 MPRG_LANE_WAIT=spin MPRG_BUILD_LANES=4
run1 n1_16c_yield6 0-15 MPRG_LANE_WAIT=yield MPRG_BUILD_LANES=6
# 1-GPU box: more lanes than the default
# for l in 8 10; do run1 n1_16c_yield$l 0-15 MPRG_BUILD_LANES=$l; done
