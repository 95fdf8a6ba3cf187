set -x
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python tools/timeline.py --reps 40 > gpurun_out/final_timeline.txt 2>&1
python tools/bench_stem.py > gpurun_out/final_stem.jsonl 2> gpurun_out/final_stem.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-encoder > gpurun_out/b_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-encoder > gpurun_out/ncu_launches2.log 2>&1
python tools/prof_step.py --steps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -o gpurun_out/final_prof_step python tools/prof_step.py --steps 1 > gpurun_out/ncu_step2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem2 -c 1 -o gpurun_out/final_prof_stem2 python tools/bench_stem.py --batches 64 --reps 3 > gpurun_out/ncu_stem2b.log 2>&1
for f in gpurun_out/final_bench.err gpurun_out/ncu_step2.log gpurun_out/ncu_stem2b.log; do tail -n 2 $f; done
