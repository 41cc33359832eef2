python tools/conv_bench.py 64 216 3 1024 1920 2 5 1 1 1 0
python tools/conv_bench.py 64 216 3 1024 1920 2 5 1 1 0 0
python tools/conv_bench.py 64 3 3 1024 1920 2 5 1 1 0 3
python tools/conv_bench.py 64 4 3 1024 1920 2 5 1 1 0 3
python tools/conv_bench.py 16 2 7 1024 1920 2 5 1 1 0 0
python tools/conv_bench.py 64 64 3 1024 1920 2 5 1 1 0 2 1
python tools/conv_bench.py 64 64 3 1024 1920 2 5 4 1 0 2 1
python tools/conv_bench.py 128 128 1 512 960 2 5
