for rep in 1 2; do
for lib in profiles/lib_old.so poccala_b200/_lib/libpoccala_b200.so; do
  echo "== $lib"
  POCCALA_B200_LIB=$PWD/$lib python profiles/time_k1.py 1000 16 2>&1 | head -1
  POCCALA_B200_LIB=$PWD/$lib python profiles/time_k1.py 12500 64 2>&1 | head -1
done
done
