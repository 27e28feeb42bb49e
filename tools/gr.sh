#!/bin/bash
# build here (nvcc cross-compiles), then run a command on the GPU box: tools/gr.sh <timeout-s> '<command>' [log-name]
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py > /dev/null
mkdir -p gpurun_out
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > "gpurun_out/${3:-call}.log" 2>&1 || true
tail -60 "gpurun_out/${3:-call}.log"
