mkdir -p gpurun_out
TCNNB_KERNEL=ws timeout 300 python tests/debug_stages.py hash3d_small 3 3 512 > gpurun_out/debug_ws.log 2>&1; echo "debug ws rc=$?"; tail -n 32 gpurun_out/debug_ws.log | head -20
TCNNB_KERNEL=ws timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_ws.log 2>&1; echo "pytest ws rc=$?"; tail -n 6 gpurun_out/pytest_ws.log
for a in 0 3; do TCNNB_KERNEL=ws TCNNB_ABLATE=$a timeout 120 python scripts/ablate.py; done 2>&1 | grep ablate | tee gpurun_out/ablate_ws.log
TCNNB_KERNEL=ws TCNNB_BINNING=0 timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee -a gpurun_out/ablate_ws.log
timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee -a gpurun_out/ablate_ws.log
