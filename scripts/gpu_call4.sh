#!/bin/bash
# Round-2 session D: host-path (e2e) A/B of batch widths / chunk sizes on c5.
tag=${1:-r2k}
mkdir -p gpurun_out
out=gpurun_out/${tag}_e2e_ab.log
: > $out
arm() { name=$1; shift; env ARM=$name "$@" python scripts/e2e_ab.py c5 12 2>/dev/null | tail -1 | tee -a $out; }
arm default X=1
arm zero15 DEODR_B200_HOST_WIDTH_ZERO=15
arm grads15 DEODR_B200_HOST_WIDTH_GRADS=15
arm grads15_1m DEODR_B200_HOST_WIDTH_GRADS=15 DEODR_B200_HOST_DOWN_CHUNK_KB=1024
arm down15 DEODR_B200_HOST_WIDTH_DOWN=15
arm down11 DEODR_B200_HOST_WIDTH_DOWN=11
arm up11 DEODR_B200_HOST_WIDTH_UP=11
arm up15 DEODR_B200_HOST_WIDTH_UP=15
arm up4 DEODR_B200_HOST_WIDTH_UP=4
arm chunk2m DEODR_B200_HOST_DOWN_CHUNK_KB=2048
arm threads24 DEODR_B200_HOST_THREADS=24 DEODR_B200_HOST_WIDTH_ZERO=23 DEODR_B200_HOST_WIDTH_GRADS=23
arm all15 DEODR_B200_HOST_WIDTH_ZERO=15 DEODR_B200_HOST_WIDTH_GRADS=15 DEODR_B200_HOST_WIDTH_DOWN=15 DEODR_B200_HOST_WIDTH_UP=15
arm default2 X=1
