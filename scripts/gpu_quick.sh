mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export TCNNB_NO_PDL=1; fi
  timeout 200 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_pdl_$v.log 2>&1; tail -n 1 gpurun_out/bench_pdl_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no_pdl=$v', d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['binning_kernels_ms'], d['roofline']['optimizer_kernel_ms'])"
done
