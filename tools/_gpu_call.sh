timeout 300 python tools/check_dense_general.py 200 0 2>&1 | tail -8
python tools/time_cheb_sizes.py > gpurun_out/time_cheb_sizes2.log 2>&1
cat gpurun_out/time_cheb_sizes2.log
