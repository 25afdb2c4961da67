timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or forward or determin" 2>&1 | tail -3
for i in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-job --no-e2e > gpurun_out/ab.json 2> gpurun_out/ab.err
python scripts/show_bench.py gpurun_out/ab.json
done
