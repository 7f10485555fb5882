R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29730
for cfg in c2 c3 c4; do
  for env in "X=1" "ISOKANN_NO_P2P=1" "ISOKANN_GRAPH=0" "ISOKANN_KOOP_FUSED=0"; do
    port=$((port+1))
    out=$(env $env timeout 120 $R --master-port $port bench.py --gpus 2 --config $cfg --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>&1 | grep -E "collapsed|\"metric\"" | head -1 | cut -c1-80)
    echo "$cfg $env -> $out"
  done
done
