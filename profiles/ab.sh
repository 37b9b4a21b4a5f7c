# A/B of two builds on the same box: profiles/ab.sh  (profiles/lib_old.so = the previous build)
for rep in 1 2; do
for lib in profiles/lib_old.so poccala_b200/_lib/libpoccala_b200.so; do
  echo "== $lib"
  POCCALA_B200_LIB=$PWD/$lib python profiles/time_k1.py 1000 16 2>&1 | head -1
  POCCALA_B200_LIB=$PWD/$lib python profiles/time_k1.py 12500 64 2>&1 | head -1
done
done
for lib in profiles/lib_old.so poccala_b200/_lib/libpoccala_b200.so; do
  echo "== bench $lib"
  POCCALA_B200_LIB=$PWD/$lib PC_BENCH_NO_CPU=1 PC_BENCH_CFG5_UTT=0 PC_BENCH_SHORT=1 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['all_ms'], d['e2e']['value'])"
done
