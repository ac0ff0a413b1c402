mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpp_shim.py -q -m gpu > gpurun_out/pytest_shim.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_shim.log
LD_LIBRARY_PATH=tiny-cuda-nn_b200 timeout 120 tests/cpp/shim_sample 300
