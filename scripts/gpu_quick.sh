mkdir -p gpurun_out
for a in 0 1 2 3; do TCNNB_ABLATE=$a timeout 120 python scripts/ablate.py; done 2>&1 | grep ablate | tee gpurun_out/ablate_ws.log
for a in 1 2; do TCNNB_BINNING=0 TCNNB_ABLATE=$a timeout 120 python scripts/ablate.py; done 2>&1 | grep ablate | tee -a gpurun_out/ablate_ws.log
