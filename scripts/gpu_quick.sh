mkdir -p gpurun_out
for v in XU_FLOOR SCATTER_3WAY NO_PARKED_WAIT; do echo $v; TCNNB_LIB=$PWD/scripts/experiments/libtcnn_b200_$v.so timeout 120 python scripts/ablate.py 2>&1 | grep ablate; done | tee gpurun_out/ablate_variants.log
echo base; timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee gpurun_out/ablate.log
echo base2; timeout 120 python scripts/ablate.py 2>&1 | grep ablate | tee -a gpurun_out/ablate.log
