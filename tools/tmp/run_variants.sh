for v in umsleep umsta umboth; do
  echo "== $v"; FME_B200_LIB=variants/libfme_$v.so FME_K2_PATH=4 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "k2_paths_bit_identical and 4" 2>&1 | tail -1
  FME_B200_LIB=variants/libfme_$v.so FME_K2_PATH=4 timeout 100 python tools/k2_by_class.py 2>&1 | head -1
done
echo "== umstaprof"; FME_B200_LIB=variants/libfme_umstaprof.so FME_K2_PATH=4 timeout 100 python tools/um_prof.py 2>&1 | tail -10
