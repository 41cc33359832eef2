# Developer A/B of two library builds on the same GPU box (under gpurun): the in-tree library vs tdvc_b200/libtdvc_b200_ab.so
set -x
for lib in "" "$PWD/tdvc_b200/libtdvc_b200_ab.so"; do
  echo "== lib: ${lib:-default}"
  for args in "64 64 3 1024 1920 2 5" "64 64 3 1024 1920 2 5 1 1 0 1 1" "4 64 3 1024 1920 2 5" "128 128 3 512 960 2 5" "64 32 7 1024 1920 2 5" "32 64 7 1024 1920 2 5" "32 16 7 1024 1920 2 5" "8 32 7 1024 1920 2 5" "128 128 1 512 960 2 5 1 1 0 0 1" "256 64 1 1024 1920 2 5" "64 128 3 1024 1920 2 5 1 2"; do
    TDVC_B200_LIB=$lib python tools/conv_bench.py $args
  done
done
