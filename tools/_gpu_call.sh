V=numgrids_b200/lib/variants
L=gpurun_out/dense_variants.log
: > $L
python tools/time_cfg4_variants.py >> $L 2>&1
for v in minb3 fk32b fk32minb3; do NUMGRIDS_B200_LIB=$PWD/$V/libng_$v.so python tools/time_cfg4_variants.py >> $L 2>&1; done
for v in ncp2 ncp2minb1; do NG_DENSE_NCP=2 NUMGRIDS_B200_LIB=$PWD/$V/libng_$v.so python tools/time_cfg4_variants.py >> $L 2>&1; done
cat $L
