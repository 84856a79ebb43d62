#!/bin/bash
# Runs on the GPU box (under gpurun): the ncu passes of /opt/skills/guides/B200_PROFILING.md for bench.py's command,
# plus the random-sector gather ceiling.  Outputs land in gpurun_out/; summaries are copied into profiles/ by hand.
#   tools/profile_round.sh <tag> [full]
set -u
TAG=${1:-r01}
FULL=${2:-}
OUT=gpurun_out
mkdir -p $OUT
KERN='regex:scan_kernel|bin_kernel|walk_kernel|order_tasks_kernel'   # the kernels of the timed region (insert_kernel builds the filter during setup)

# 1. practical ceiling of random 32-byte-sector reads over a 4 GiB buffer
if [ -x tools/gather_bench ]; then
	tools/gather_bench 4 0 0 > $OUT/gather_${TAG}.jsonl 2>&1
fi
if [ -x tools/gather_modes ]; then
	tools/gather_modes 4 > $OUT/gather_modes_${TAG}.jsonl 2>&1
fi

# 2. launch list of the bench command (our kernels only; cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv \
	--log-file $OUT/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1

# 3. full captures: the first bin / probe pair of the scan stage, and the first (largest) walker launch
if [ -n "$FULL" ]; then
	ncu --set full --clock-control none --import-source on -k 'regex:bin_kernel' -c 2 -f \
		-o $OUT/prof_scan_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/bench_under_ncu_full_scan_${TAG}.log 2>&1
	ncu --set full --clock-control none --import-source on -k 'regex:walk_kernel' -c 1 -f \
		-o $OUT/prof_walk_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/bench_under_ncu_full_walk_${TAG}.log 2>&1
fi
ls -la $OUT
