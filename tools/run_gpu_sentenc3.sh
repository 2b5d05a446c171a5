mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sentenc.py -q -m gpu > gpurun_out/pytest_sentenc.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_sentenc.log
