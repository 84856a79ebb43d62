#!/bin/bash
# Runs on the GPU box (under gpurun): the ncu passes of /opt/skills/guides/B200_PROFILING.md for bench.py's command.
# Outputs land in gpurun_out/; summaries are made here with tools/ncu_summary.py and copied into profiles/ by hand.
#   tools/profile_round.sh <tag> [full]
set -u
TAG=${1:-r02}
FULL=${2:-}
OUT=gpurun_out
mkdir -p $OUT
# the kernels of the timed region (insert_kernel / occupancy_kernel build the filter during setup)
KERN='regex:scan_kernel|bin_kernel|heads_kernel|presite|snv_dense|walk_kernel|order_tasks_kernel|compact_events_kernel|fetch_host_kernel'
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"

# 1. launch list of the bench command (our kernels only; cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 2000 --csv \
	--log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_bench_under_ncu.log 2>&1

# 2. full captures, one invocation per kernel family (the launch skip counts put the capture in the first timed-style call)
if [ -n "$FULL" ]; then
	ncu --set full --clock-control none --import-source on -k 'regex:bin_kernel' -c 4 -f \
		-o $OUT/${TAG}_prof_scan $BENCH > $OUT/${TAG}_ncu_scan.log 2>&1
	ncu --set full --clock-control none --import-source on -k 'regex:^presite_dense_kernel' -c 2 -f \
		-o $OUT/${TAG}_prof_dense $BENCH > $OUT/${TAG}_ncu_dense.log 2>&1
	ncu --set full --clock-control none --import-source on -k 'regex:^presite_kernel' -c 1 -f \
		-o $OUT/${TAG}_prof_second $BENCH > $OUT/${TAG}_ncu_second.log 2>&1
	ncu --set full --clock-control none --import-source on -k 'regex:walk_kernel' -c 1 -f \
		-o $OUT/${TAG}_prof_walk $BENCH > $OUT/${TAG}_ncu_walk.log 2>&1
	ncu --set full --clock-control none --import-source on -k 'regex:snv_dense_kernel' -c 1 -f \
		-o $OUT/${TAG}_prof_snv python bench.py --workload tiny_snv --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_snv.log 2>&1
fi
# gpurun brings back at most 64 MiB: keep the per-kernel summary and the raw metric pages, not the reports
if [ -n "$FULL" ]; then
	python tools/ncu_summary.py $OUT/${TAG}_prof_scan.ncu-rep $OUT/${TAG}_prof_dense.ncu-rep $OUT/${TAG}_prof_second.ncu-rep \
		$OUT/${TAG}_prof_walk.ncu-rep $OUT/${TAG}_prof_snv.ncu-rep > $OUT/${TAG}_ncu_full_kernels_summary.csv
	for r in scan dense second walk snv; do
		ncu -i $OUT/${TAG}_prof_$r.ncu-rep --page raw --csv --print-units base > $OUT/${TAG}_prof_${r}_raw.csv 2>/dev/null
		rm -f $OUT/${TAG}_prof_$r.ncu-rep
	done
fi
ls -la $OUT | grep ${TAG}_
