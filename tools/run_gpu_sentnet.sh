mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sentenc.py -q -m gpu -x -k "sentence_vector_net" > gpurun_out/pytest_sentnet.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_sentnet.log
