for v in g8_s20 g16_s0 g16_s20 g32_s0 g32_s40; do
echo "== $v"; COATI_GPU_LIB=$PWD/tools/gpu/ab_$v.so timeout 300 python tools/wave_exp.py 2>&1 | grep -E "R (4|10) "
done
