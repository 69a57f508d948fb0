python -m pytest tests/test_layerwise_gpu.py tests/test_bench_shape_gpu.py tests/test_pruning_gpu.py tests/test_data_gpu.py -m gpu -q -rP 2>&1 | grep -v "^$" > gpurun_out/t_new.log
tail -3 gpurun_out/t_new.log
for c in uc2_train m3p_train uc2_sft70 eval_uc2 eval_m3p; do python bench.py --config $c --no-cpu-baseline --steps 8 > gpurun_out/bench_$c.log 2>&1; tail -c 300 gpurun_out/bench_$c.log; done
for c in uc2_train m3p_train eval_uc2; do python bench.py --impl torch_gpu --config $c --steps 8 > gpurun_out/torch_$c.log 2>&1; tail -c 300 gpurun_out/torch_$c.log; done
python bench.py --config uc2_train --features bf16 --no-cpu-baseline --steps 8 > gpurun_out/bench_uc2_bf16feat.log 2>&1
