#!/bin/bash
# One pass of the round's evidence on a B200 box (run under gpurun from the repo root); everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash profiles/run_profiles.sh r1'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# launch list of one steady-state step (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/${tag}_launches.csv \
    python profiles/profile_step.py > /dev/null 2>&1
# ncu --set full of the dominant kernels, one launch each
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o $out/${tag}_gemm_qkv python profiles/gemm_one.py 2304 768 bf16 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o $out/${tag}_gemm_fc1 python profiles/gemm_one.py 3072 768 gelu 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o $out/${tag}_gemm_fc2 python profiles/gemm_one.py 768 3072 res 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_global -s 3 -c 1 -o $out/${tag}_attn_global python profiles/attn_one.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_window -s 3 -c 1 -o $out/${tag}_attn_window python profiles/attn_win_one.py > /dev/null 2>&1
python profiles/gemm_micro.py 2 > $out/${tag}_gemm_micro.txt 2>&1
python profiles/attn_one.py > $out/${tag}_attn_micro.txt 2>&1
python profiles/attn_win_one.py >> $out/${tag}_attn_micro.txt 2>&1
cat $out/${tag}_pytest_gpu.log
cut -c1-400 $out/${tag}_bench_n1.json
cut -c1-400 $out/${tag}_bench_reference.json
ls -la $out | tail -20
