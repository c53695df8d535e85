#!/bin/bash
# usage (under gpurun): bash tools/exp_lat_sweep.sh <tag>   -- wide / split vs latency kernels over stream counts
tag=${1:-r2}
mkdir -p gpurun_out
out=gpurun_out/${tag}_ops_latency.jsonl
: > $out
for lat in 0 1; do
  ISSCABAC_LAT=$lat python tools/exp_ops_latency.py --streams 32,1024,4096,8192,16384,32768,65536 2>&1 | sed "s/\"enc_split\"/\"lat\": \"$lat\", \"enc_split\"/" >> $out
done
cat $out
