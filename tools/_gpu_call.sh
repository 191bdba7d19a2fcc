timeout 300 python tools/check_fd_stream.py 400 8 > gpurun_out/check_fd_stream.log 2>&1; echo "rc=$?" >> gpurun_out/check_fd_stream.log
tail -4 gpurun_out/check_fd_stream.log
timeout 200 python tools/time_fd.py > gpurun_out/time_fd11.log 2>&1
grep "axis 2\|8192) order . axis 1" gpurun_out/time_fd11.log
