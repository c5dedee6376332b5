run() { echo "## $1 $2"; BM25_B200_LIB=$PWD/build/ab/$1.so python scripts/quick_gpu.py --workloads B,10M,C --configs $2 | grep -v '^#'; }
run head default
run now default
run head default
run now default
