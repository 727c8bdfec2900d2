#!/bin/bash
# dev helper: rebuild libnrw.so, then run a command on the B200 box (never ships a stale .so)
set -e
cd "$(dirname "$0")"
./neuralrecon-w_b200/build.sh 2>&1 | grep -v deprecated | tail -3
T=${GPU_TIMEOUT:-900}
exec /usr/local/graft/bin/gpurun --timeout $T -- "$@"
