#!/bin/bash
# Round-2 session E: host-path placement (which NUMA node the caller runs on) and batch widths, interleaved repeats.
tag=${1:-r2l}
mkdir -p gpurun_out
out=gpurun_out/${tag}_e2e_ab.log
: > $out
nvidia-smi topo -m 2>&1 | head -8 | tee -a $out
lscpu | grep -i -e "numa" -e "model name" -e "^CPU(s)" | tee -a $out
cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0 | tr 'A-Z' 'a-z' | sed 's/^0000//')/numa_node 2>/dev/null | tee -a $out
n0=$(cat /sys/devices/system/node/node0/cpulist); n1=$(cat /sys/devices/system/node/node1/cpulist 2>/dev/null)
echo "node0 $n0 node1 $n1" | tee -a $out
arm() { name=$1; shift; env ARM=$name "$@" 2>/dev/null | tail -1 | tee -a $out; }
for rep in 1 2 3; do
arm default python scripts/e2e_ab.py c5 10
arm node0 taskset -c $n0 python scripts/e2e_ab.py c5 10
[ -n "$n1" ] && arm node1 taskset -c $n1 python scripts/e2e_ab.py c5 10
arm up15 env DEODR_B200_HOST_WIDTH_UP=15 python scripts/e2e_ab.py c5 10
arm all15 env DEODR_B200_HOST_WIDTH_ZERO=15 DEODR_B200_HOST_WIDTH_GRADS=15 DEODR_B200_HOST_WIDTH_DOWN=15 DEODR_B200_HOST_WIDTH_UP=15 python scripts/e2e_ab.py c5 10
done
