#!/bin/bash
# helper run on the GPU box: $1 = job name (see the cases)
set -u
out=gpurun_out
mkdir -p $out
launch_list() {   # $1 = CONFIG, $2 = tag
  CONFIG=$1 python tools/profile_step.py > $out/prof_plain_$2.log 2>&1 || { echo "profile_step $1 failed"; tail -5 $out/prof_plain_$2.log; return; }
  CONFIG=$1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file $out/$2_launches.csv python tools/profile_step.py > $out/ncu_list_$2.log 2>&1
  python tools/agg_launches.py $out/$2_launches.csv 40 > $out/$2_launches.txt 2>&1
  head -25 $out/$2_launches.txt
}
case "$1" in
  opt1)
    python -m pytest tests/test_ops2_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_m3p_gpu.py -m gpu -q -x 2>&1 | tail -5
    python bench.py --no-cpu-baseline --steps 10 > $out/bench_opt1.log 2>&1; tail -c 1500 $out/bench_opt1.log | head -c 600; echo
    launch_list eval_uc2 r02_eval_uc2
    launch_list m3p_train r02_m3p_train
    launch_list uc2_train r02_uc2_train_a
    ;;
  wide1)
    python -m pytest tests/test_ops_gpu.py tests/test_ops2_gpu.py tests/test_m3p_gpu.py tests/test_model_gpu.py tests/test_layerwise_gpu.py -m gpu -q -x 2>&1 | tail -8
    for c in uc2_train eval_uc2 eval_m3p m3p_train; do python bench.py --config $c --no-cpu-baseline --steps 8 > $out/bench_w1_$c.log 2>&1; tail -c 1200 $out/bench_w1_$c.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' | head -4 | tr '\n' ' '; echo $c; done
    launch_list eval_uc2 r02_eval_uc2_b
    ;;
  wide2)
    python -m pytest tests/test_ops_gpu.py tests/test_m3p_gpu.py tests/test_layerwise_gpu.py -m gpu -q -x -k "attention or m3p or eval" 2>&1 | tail -8
    for c in eval_uc2 m3p_train; do python bench.py --config $c --no-cpu-baseline --steps 8 > $out/bench_w2_$c.log 2>&1; tail -c 1200 $out/bench_w2_$c.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' | head -4 | tr '\n' ' '; echo $c; done
    launch_list eval_uc2 r02_eval_uc2_c
    ;;
  ncu1)
    python -m pytest tests/test_ops_gpu.py tests/test_m3p_gpu.py tests/test_model_gpu.py tests/test_layerwise_gpu.py -m gpu -q -x 2>&1 | tail -3
    for c in uc2_train eval_uc2 m3p_train; do python bench.py --config $c --no-cpu-baseline --steps 8 > $out/bench_n1_$c.log 2>&1; tail -c 1500 $out/bench_n1_$c.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -2 | tr '\n' ' '; echo $c; done
    python tools/profile_step.py > $out/prof_plain_ncu1.log 2>&1 || { echo "plain profile_step failed"; exit 1; }
    cap() {  # $1 = tag, $2 = CONFIG, $3 = demangled-name regex, $4 = skip
      CONFIG=$2 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$3" -s $4 -c 1 \
          -o /tmp/full_$1 -f python tools/profile_step.py > $out/ncu_full_$1.log 2>&1
      ncu -i /tmp/full_$1.ncu-rep --page details > $out/r02_ncu_full_$1.txt 2>/dev/null
      python tools/ncu_summary.py /tmp/full_$1.ncu-rep > $out/r02_ncu_summary_$1.txt 2>/dev/null
      ncu -i /tmp/full_$1.ncu-rep --page source --csv > /tmp/src_$1.csv 2>/dev/null
      python tools/ncu_source_top.py /tmp/src_$1.csv 45 > $out/r02_ncu_stalls_$1.txt 2>/dev/null
      head -4 $out/r02_ncu_summary_$1.txt
    }
    cap gemm_wgrad uc2_train 'gemm_tc_kernel<.int.256, .bool.0, .bool.0, .int.2, .int.0>' 6
    cap gemm_gelu_pair uc2_train 'gemm_tc_kernel<.int.256, .bool.1, .bool.1, .int.2, .int.3>' 4
    cap gemm_mul_aux uc2_train 'gemm_tc_kernel<.int.256, .bool.1, .bool.0, .int.2, .int.4>' 4
    cap gemm_resid_ffn2 uc2_train 'gemm_tc_kernel<.int.256, .bool.1, .bool.1, .int.2, .int.2>' 9
    cap attn_wide eval_uc2 'attn_pipe_fwd_wide_kernel' 4
    cap attn_bwd uc2_train 'attn_pipe_bwd_kernel' 3
    launch_list eval_uc2 r02_eval_uc2_d
    ;;
  relax1)
    python -m pytest tests/test_ops_gpu.py tests/test_ops2_gpu.py tests/test_gemm_fast_gpu.py tests/test_model_gpu.py tests/test_m3p_gpu.py -m gpu -q -x 2>&1 | tail -3
    for c in uc2_train eval_uc2 m3p_train; do python bench.py --config $c --no-cpu-baseline --steps 10 > $out/bench_r1_$c.log 2>&1; tail -c 1500 $out/bench_r1_$c.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -2 | tr '\n' ' '; echo $c; done
    launch_list uc2_train r02_uc2_train_b
    ;;
esac
