# the headline evidence only (one gpurun call): default bench line, reference arm, in-graph timelines incl. the zero-compute floor,
# forward ablation, ncu launch list + full capture of the step's kernels
set -x
R=gpurun_out
python bench.py > $R/r2_bench_n1.json 2> $R/r2_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $R/r2_bench_reference_arm.json 2> $R/r2_ref.err
python tools/timeline.py --reps 40 > $R/r2_timeline_b16.txt 2>&1
QW_DBG_FWD=7 QW_DBG_GY=1 python tools/timeline.py --reps 40 > $R/r2_timeline_b16_zero_compute.txt 2>&1
python bench.py --stem-forward split --no-cpu-baseline --no-encoder > $R/r2_bench_n1_split_forward.json 2> /dev/null
python tools/dbg_fwd.py > $R/r2_dbg_fwd.txt 2>&1
python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-encoder > $R/b_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $R/r2_launches_bench_b16.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-encoder > $R/ncu_launches2.log 2>&1
python tools/prof_step.py --steps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -o $R/r2_prof_step python tools/prof_step.py --steps 1 > $R/ncu_step2.log 2>&1
tail -n 2 $R/r2_bench_n1.err $R/ncu_step2.log
