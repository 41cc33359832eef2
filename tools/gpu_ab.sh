B=$PWD/tdvc_b200/libtdvc_b200_B.so
for i in 1 2 3; do
python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 2 1 | sed 's/^/A /'
TDVC_B200_LIB=$B python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 2 1 | sed 's/^/B /'
python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 1 0 | sed 's/^/A /'
TDVC_B200_LIB=$B python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 1 0 | sed 's/^/B /'
done
