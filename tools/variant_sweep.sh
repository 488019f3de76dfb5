for v in g1b3 g2b3 g4b3 g1b2 g2b2 g4b2; do
echo "variant=$v"
APOLLO_B200_LIB=build/libmsda_$v.so timeout 200 python tools/fused_bench.py 2>&1 | tail -1
done
