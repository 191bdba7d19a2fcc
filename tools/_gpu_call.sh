for u in 8; do echo "U=$u"; NG_FD_ROW_U=$u timeout 200 python tools/time_fd.py 2>&1 | grep "axis 2\|8192) order . axis 1" | grep -v "256, 256"; done > gpurun_out/time_fd10.log 2>&1
cat gpurun_out/time_fd10.log
