timeout 300 python tools/check_fft_strided.py 120 0 2>&1 | tail -6
python tools/time_fft_axes.py > gpurun_out/time_fft_axes2.log 2>&1
cat gpurun_out/time_fft_axes2.log
