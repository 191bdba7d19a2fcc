V=numgrids_b200/lib/variants
L=gpurun_out/fft_tail_variants.log
: > $L
for v in ffttb1 ffttb2 default ffttb8; do
  if [ $v = default ]; then python tools/time_cfg4_variants.py >> $L 2>&1; else NUMGRIDS_B200_LIB=$PWD/$V/libng_$v.so python tools/time_cfg4_variants.py >> $L 2>&1; fi
done
cut -c1-120 $L
