#!/bin/bash
# Runs on the GPU box: plain bench first (exit code checked), then the ncu launch list of one profiled step and
# --set full captures of the dominant kernels; only text / csv summaries are kept in gpurun_out/.
set -u
tag=${1:-r01}
out=gpurun_out
python bench.py --steps 10 --warmup 3 > $out/bench_${tag}.json 2> $out/bench_${tag}.err || { echo "bench failed"; tail -5 $out/bench_${tag}.err; exit 1; }
python tools/show_bench.py $out/bench_${tag}.json
python tools/profile_step.py > $out/prof_plain.log 2>&1 || { echo "profile_step failed"; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $out/${tag}_launches_step_b256.csv python tools/profile_step.py > $out/ncu_list.log 2>&1
tail -1 $out/ncu_list.log
for spec in "gemm_qkv_plain:gemm_tc:45" "gemm_resid:gemm_tc:20" "attn_bwd:attn_pipe_bwd:3" "attn_fwd:attn_pipe_fwd:3" "adamw:adamw_kernel:0"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%%:*}; skip=${rest#*:}
  ncu --profile-from-start off --set full --clock-control none -k regex:$kern -s $skip -c 1 -o /tmp/full_$name -f python tools/profile_step.py > $out/ncu_full_$name.log 2>&1
  ncu -i /tmp/full_$name.ncu-rep --page details > $out/${tag}_ncu_full_$name.txt 2>/dev/null
  ncu -i /tmp/full_$name.ncu-rep --page raw --csv > /tmp/raw_$name.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/full_$name.ncu-rep > $out/${tag}_ncu_summary_$name.txt 2>/dev/null
  head -3 $out/${tag}_ncu_summary_$name.txt
done
