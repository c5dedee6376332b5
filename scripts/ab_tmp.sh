run() { echo "## $1 $2"; BM25_B200_LIB=$PWD/build/ab/$1.so python scripts/quick_gpu.py --workloads B,10M,E --configs $2 | grep -v '^#'; }
BM25_B200_LIB=$PWD/build/ab/v6.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run b256x3 default
run v6 default
