B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline"
V=nerfies_b200/_variants
echo "prod:"; timeout 200 $B 2>/dev/null | python tools/benchline.py
for v in y1 x1 x2 x3 x4 x5 x6; do
  echo "$v:"; NFB_LIB_PATH=$V/libnfb_$v.so timeout 200 $B 2>/dev/null | python tools/benchline.py
done
timeout 100 python -m pytest tests/test_parity_gpu.py tests/test_tc_selftest_gpu.py -x -q 2>&1 | tail -2
