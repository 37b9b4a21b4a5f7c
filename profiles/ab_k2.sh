for lib in profiles/lib_old.so poccala_b200/_lib/libpoccala_b200.so; do
  echo "== $lib"
  POCCALA_B200_LIB=$PWD/$lib python profiles/exp_k2.py 2>&1 | grep "k2_kernel 1"
done
for lib in profiles/lib_old.so poccala_b200/_lib/libpoccala_b200.so; do
  echo "== bench $lib"
  POCCALA_B200_LIB=$PWD/$lib PC_BENCH_NO_CPU=1 PC_BENCH_CFG5_UTT=0 PC_BENCH_SHORT=1 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['all_ms'], d['e2e']['value'])"
done
