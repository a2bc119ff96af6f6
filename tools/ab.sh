for v in "" ab_inline.so ab_mb3.so ab_mb5.so ab_mb6.so ab_mb8.so; do
  echo "== variant ${v:-default}"
  POLARIS_CUDA_LIB=${v:+$PWD/$v} timeout 200 python bench.py --steps 2 --warmup 1 --spp 64 --no-cpu 2>&1 | grep -E "timed|kernel classes" | sed -e 's/"alg_GBps": [0-9.]*//g' | cut -c1-420
done
