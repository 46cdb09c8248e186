for cfg in "16 16" "8 64" "4 32" "8 128" "16 64"; do set -- $cfg; CFL_BN_PPT_SMALL=$1 CFL_BN_PPT_LARGE=$2 timeout 100 python scripts/bench_bn.py >> gpurun_out/bench_bn.jsonl 2>> gpurun_out/bench_bn.err; done
cat gpurun_out/bench_bn.jsonl
timeout 400 python -m pytest tests/test_gpu_optim.py tests/test_gpu_tower_ops.py tests/test_gpu_gemm.py tests/test_gpu_towers.py tests/test_gpu_step_parity.py -q -x 2>&1 | tail -15 > gpurun_out/subset_tests.log; tail -4 gpurun_out/subset_tests.log
timeout 100 python scripts/sweep_wgrad.py > gpurun_out/sweep_wgrad2.jsonl 2>/dev/null; cat gpurun_out/sweep_wgrad2.jsonl | cut -c1-200
