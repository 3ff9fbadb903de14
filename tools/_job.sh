timeout 900 python -m pytest tests/test_gpu_at_size.py -m gpu -x -q -k "random_parameter" > gpurun_out/r02ae_pytest.log 2>&1; tail -12 gpurun_out/r02ae_pytest.log | cut -c1-400
python tools/stream_step.py 2>&1 | head -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r02ae_pytest2.log 2>&1; tail -3 gpurun_out/r02ae_pytest2.log
