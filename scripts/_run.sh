for w in 1 0 1; do echo "== pool 16 wide $w"; DEODR_B200_HOST_WIDE=$w python scripts/e2e_trace.py c5 2>&1 | grep "iter 3\|host path\|DMA\|stage\|mirror" ; done
