run() { echo "## $1 $2"; BM25_B200_LIB=$PWD/build/ab/$1.so python scripts/quick_gpu.py --workloads B,10M,C,E --configs $2 | grep -v '^#'; }
BM25_B200_LIB=$PWD/build/ab/bulkcold.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -1
run head default
run bulkcold default
run head default
run bulkcold default
