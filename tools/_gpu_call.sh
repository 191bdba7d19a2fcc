python tools/run_fd_fused.py 512 2 post > gpurun_out/plain_fdf.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'fd_row' -s 2 -c 1 -f -o gpurun_out/prof_fd_row_fused python tools/run_fd_fused.py 512 2 post > gpurun_out/ncu_fdf.log 2>&1
echo done
