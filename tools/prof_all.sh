# round-2 evidence run (one gpurun call): final bench lines, timelines, sweeps, ncu launch list + full capture of the step's kernels
set -x
R=gpurun_out
python bench.py > $R/r2_bench_n1.json 2> $R/r2_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $R/r2_bench_reference_arm.json 2> $R/r2_ref.err
python tools/timeline.py --reps 40 > $R/r2_timeline_b16.txt 2>&1
QW_DBG_FWD=7 QW_DBG_GY=1 python tools/timeline.py --reps 40 > $R/r2_timeline_b16_zero_compute.txt 2>&1
python tools/bench_stem.py > $R/r2_stem_fused_vs_unfused.jsonl 2> $R/r2_stem.err
python tools/sweep_circuit.py > $R/r2_config4_circuit_sweep.jsonl 2> $R/r2_sweep4.err
python tools/sweep_circuit.py --embedding angle > $R/r2_config4_circuit_sweep_angle.jsonl 2>> $R/r2_sweep4.err
python tools/sweep_logmel.py > $R/r2_config5_logmel_stem_sweep.jsonl 2> $R/r2_sweep5.err
python tools/prof_act.py > $R/r2_gelu_fused_kernels.txt 2>&1
python tools/bench_general.py > $R/r2_general_path_stem_b16.jsonl 2> $R/r2_general.err
python tools/prof_general.py --q 6 8 10 > $R/r2_general_path_kernels_b16.jsonl 2>> $R/r2_general.err
python tools/prof_general.py --q 4 --embedding angle >> $R/r2_general_path_kernels_b16.jsonl 2>> $R/r2_general.err
# the probe binaries are not shipped to the GPU box (.gpurunignore): build them there
nvcc -cudart shared -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/probe/mma_probe.cu -o tools/probe/mma_probe 2> $R/probe_build.err && \
./tools/probe/mma_probe > $R/r2_mma_probe.txt 2>&1
nvcc -cudart shared -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/probe/stream_probe.cu -o tools/probe/stream_probe -lcuda 2>> $R/probe_build.err && \
tools/probe/run_stream_probe.sh > $R/r2_stream_probe.txt 2>&1
python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-encoder > $R/b_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $R/r2_launches_bench_b16.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-encoder > $R/ncu_launches2.log 2>&1
python tools/prof_step.py --steps 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -o $R/r2_prof_step python tools/prof_step.py --steps 1 > $R/ncu_step2.log 2>&1
python tools/prof_circuit.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wcirc -c 3 -o $R/r2_prof_circuit_q10 python tools/prof_circuit.py > $R/ncu_circ.log 2>&1
python tools/prof_logmel.py 64 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:logmel -c 3 -o $R/r2_prof_logmel_b64 python tools/prof_logmel.py 64 > $R/ncu_logmel.log 2>&1
for f in $R/r2_bench_n1.err $R/ncu_step2.log $R/ncu_circ.log $R/ncu_launches2.log; do tail -n 2 $f; done
