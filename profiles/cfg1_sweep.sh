for m in 1 2 3 4 6 8; do
  echo "UPMIX_DEC_RUN_MIN=$m"
  UPMIX_DEC_RUN_MIN=$m python profiles/config_bench.py 2>&1 | grep "10 s" | cut -c1-36,95-
done
