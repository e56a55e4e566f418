#!/bin/bash
# Frame-shard measurements on an N-GPU box:  scripts/gpu_shard_suite.sh N [tests]
# writes gpurun_out/shard_n${N}_*.json (one bench line each) and, with "tests", the sharded-vs-unsharded parity log.
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, extra args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 2 --warmup 3 "$@" > $OUT/shard_n${N}_${name}.json 2> $OUT/shard_n${N}_${name}.err
  echo "== $name rc=$?"; tail -c 1200 $OUT/shard_n${N}_${name}.json; echo
}
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_frame_shard_gpu.py -q -s -k "$N" > $OUT/shard_n${N}_tests.log 2>&1
  echo "tests rc=$?"; grep "\[shard\]\|passed\|failed" $OUT/shard_n${N}_tests.log | tail -20
fi
run fs125 --mode frame_shard --frames 125
run fscfg125 --mode frame_shard_cfg --frames 125
run fs24 --mode frame_shard --frames 24
run fscfg24 --mode frame_shard_cfg --frames 24
