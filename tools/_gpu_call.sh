timeout 200 python tools/check_fd_stream.py 300 13 2>&1 | tail -1
timeout 100 python tools/time_fd_fused.py 512 2>&1 | grep "axis 2"
