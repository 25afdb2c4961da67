# Evidence capture of a round (run on the GPU box through gpurun): default bench line, ncu launch list of the eager step,
# ncu --set full of the heavy kernels.  Summaries: scripts/ncu_summary.py -> profiles/.
set -x
timeout 900 python bench.py > gpurun_out/r2_pair_bench_n1.json 2> gpurun_out/r2_pair_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-job --no-e2e > gpurun_out/r2_pair_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_pair_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-job --no-e2e > gpurun_out/r2_pair_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_pair|segment_reduce|gemm_pair" -s 57 -c 10 -o gpurun_out/r2_pair_kernels -f python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-job --no-e2e > gpurun_out/r2_pair_ncu_full.log 2>&1
ls -la gpurun_out/r2_pair_*
