timeout 200 python tools/check_fd_stream.py 300 12 2>&1 | tail -1
timeout 100 python tools/time_fd.py 2>&1 | grep "1024, 1024, 1024" | grep "log\|order 1 axis 0"
