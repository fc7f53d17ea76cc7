#!/bin/bash
# Round-2 session F: host-path crew size for the DRAM-bound adjoint batch, interleaved repeats.
tag=${1:-r2m}
mkdir -p gpurun_out
out=gpurun_out/${tag}_e2e_ab.log
: > $out
arm() { name=$1; shift; env ARM=$name "$@" 2>/dev/null | tail -1 | tee -a $out; }
W="DEODR_B200_HOST_WIDTH_UP=15 DEODR_B200_HOST_WIDTH_GRADS=15"
for rep in 1 2 3; do
arm A_up15_grads15 env $W python scripts/e2e_ab.py c5 10
arm B_threads32 env $W DEODR_B200_HOST_THREADS=32 python scripts/e2e_ab.py c5 10
arm C_threads24 env $W DEODR_B200_HOST_THREADS=24 python scripts/e2e_ab.py c5 10
arm D_threads32_up23 env DEODR_B200_HOST_WIDTH_UP=23 DEODR_B200_HOST_WIDTH_GRADS=15 DEODR_B200_HOST_THREADS=32 python scripts/e2e_ab.py c5 10
arm E_threads48 env $W DEODR_B200_HOST_THREADS=48 python scripts/e2e_ab.py c5 10
done
