#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): GPU parity tests, default bench line, ncu launch lists of one inference step
# and one training step, the training-step timings of configs 2 / 4.
#   gpurun --timeout 2400 -- 'bash profiles/run_profiles_r2.sh [tag]'
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python profiles/profile_step.py > gpurun_out/${TAG}_launches.log 2>&1
echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
head -24 gpurun_out/${TAG}_launches_summary.txt
# training step (BASELINE config 4's step on ViT-B 8x1024^2, ViT-H 32x512^2 and ViT-H 32x1024^2)
{ python profiles/train_step.py; python profiles/train_step.py vit_h 32 512 4; python profiles/train_step.py vit_h 32 1024 4; } 2>/dev/null | grep workload > gpurun_out/${TAG}_train_step.json
cat gpurun_out/${TAG}_train_step.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/${TAG}_launches_train.csv \
    python profiles/train_step.py --profile > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/${TAG}_launches_train.csv > gpurun_out/${TAG}_launches_train_summary.txt 2>&1
head -24 gpurun_out/${TAG}_launches_train_summary.txt
