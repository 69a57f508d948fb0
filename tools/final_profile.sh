#!/bin/bash
# Final single-GPU evidence run of a round (under gpurun): every bench config (ours and the stock-PyTorch arm), the kernel-vs-cuBLAS
# table, then - only after the plain runs exited 0 - the ncu launch lists and --set full captures.  Text / csv only in gpurun_out/.
set -u
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
fail=0
for c in uc2_train m3p_train uc2_sft41 uc2_sft70 eval_uc2 eval_m3p; do
  extra=""; [ $c != uc2_train ] && extra="--no-cpu-baseline"
  python bench.py --config $c --steps 10 --warmup 3 $extra > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err || { echo "bench $c failed"; tail -3 $out/${tag}_bench_$c.err; fail=1; }
done
python bench.py --config uc2_train --features bf16 --no-cpu-baseline --steps 10 > $out/${tag}_bench_uc2_train_bf16feat.json 2>/dev/null
for c in uc2_train m3p_train eval_uc2 eval_m3p; do
  python bench.py --impl torch_gpu --config $c --steps 10 --warmup 3 > $out/${tag}_torchgpu_$c.json 2> $out/${tag}_torchgpu_$c.err || echo "torch_gpu $c failed"
done
python tools/show_bench.py $out/${tag}_bench_*.json $out/${tag}_torchgpu_*.json
python tools/bench_gemm.py > $out/${tag}_gemm_vs_cublas.txt 2>&1; tail -5 $out/${tag}_gemm_vs_cublas.txt
python tools/bench_attn.py > $out/${tag}_attention_micro.txt 2>&1; S=140 B=256 SMIN=106 python tools/bench_attn.py 2>/dev/null | grep "pipelined  " >> $out/${tag}_attention_micro.txt
[ $fail = 1 ] && exit 1
launch_list() {   # $1 = CONFIG
  CONFIG=$1 python tools/profile_step.py > $out/prof_plain_$1.log 2>&1 || { echo "profile_step $1 failed"; return; }
  CONFIG=$1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file $out/${tag}_launches_$1.csv python tools/profile_step.py > $out/ncu_list_$1.log 2>&1
  python tools/agg_launches.py $out/${tag}_launches_$1.csv 45 > $out/${tag}_launches_$1.txt 2>&1
  head -3 $out/${tag}_launches_$1.txt
}
launch_list uc2_train; launch_list m3p_train; launch_list eval_uc2
cap() {  # $1 = name, $2 = CONFIG, $3 = demangled-name regex, $4 = launches to skip
  CONFIG=$2 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$3" -s $4 -c 1 \
      -o /tmp/full_$1 -f python tools/profile_step.py > $out/ncu_full_$1.log 2>&1
  ncu -i /tmp/full_$1.ncu-rep --page details > $out/${tag}_ncu_full_$1.txt 2>/dev/null
  python tools/ncu_summary.py /tmp/full_$1.ncu-rep > $out/${tag}_ncu_summary_$1.txt 2>/dev/null
  ncu -i /tmp/full_$1.ncu-rep --page source --csv > /tmp/src_$1.csv 2>/dev/null
  python tools/ncu_source_top.py /tmp/src_$1.csv 40 > $out/${tag}_ncu_stalls_$1.txt 2>/dev/null
  head -2 $out/${tag}_ncu_summary_$1.txt
}
cap gemm_wgrad uc2_train 'gemm_tc_kernel<.int.256, .bool.0, .bool.0, .int.2, .int.0>' 6
cap gemm_gelu_pair uc2_train 'gemm_tc_kernel<.int.256, .bool.1, .bool.1, .int.2, .int.3>' 4
cap gemm_mul_aux uc2_train 'gemm_tc_kernel<.int.256, .bool.1, .bool.0, .int.2, .int.4>' 4
cap gemm_resid_ffn2 uc2_train 'gemm_tc_kernel<.int.256, .bool.1, .bool.1, .int.2, .int.2>' 9
cap gemm_qkv_plain uc2_train 'gemm_tc_kernel<.int.256, .bool.1, .bool.1, .int.2, .int.1>' 4
cap attn_fwd uc2_train 'attn_pipe_fwd_kernel' 3
cap attn_bwd uc2_train 'attn_pipe_bwd_kernel' 3
cap attn_fwd_wide eval_uc2 'attn_pipe_fwd_wide_kernel' 4
cap adamw uc2_train 'adamw_kernel' 0
cap ln_fwd uc2_train 'ln_fwd_vec_kernel' 10
