#!/bin/bash
# Round-2 ncu evidence, run on a B200 under gpurun (one GPU):  bash tools/capture_profiles.sh
# Every command is first run plainly (must exit 0), then under ncu; the .ncu-rep files are summarised on the box
# (tools/ncu_summary.py) because gpurun_out/ is limited to 64 MiB; only the C2 report travels back.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -k regex:trace_kernel"
prof() {  # name, launches, command...
    local name=$1 count=$2; shift 2
    "$@" > $OUT/plain_$name.log 2>&1 || { echo "plain run of $name failed"; tail -5 $OUT/plain_$name.log; return; }
    $NCU -c $count -o $OUT/$name -f "$@" > $OUT/ncu_$name.log 2>&1
    for row in $(seq 0 $((count - 1))); do
        python tools/ncu_summary.py $OUT/$name.ncu-rep $row $OUT/${name}_launch$row.json > /dev/null 2>&1
    done
    cat $OUT/plain_$name.log
}
prof r2_c2_final_1500 2 python tools/render_once.py C2
prof r2_c1_final_500 2 python tools/render_once.py C1
prof r2_c3_final_640 1 python tools/render_once.py C3 640 640 2000
prof r2_c5_final_1024 1 python tools/render_once.py C5 1024 1024
rm -f $OUT/r2_c1_final_500.ncu-rep $OUT/r2_c3_final_640.ncu-rep $OUT/r2_c5_final_1024.ncu-rep
BENCH="python bench.py --steps 2 --warmup 1 --per-scene-budget-s 0 --no-cpu-baseline"
$BENCH > $OUT/plain_bench.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_launches_c2_1500.csv $BENCH > $OUT/ncu_bench.log 2>&1
tail -1 $OUT/plain_bench.log | cut -c1-400
ls -la $OUT | tail -20
