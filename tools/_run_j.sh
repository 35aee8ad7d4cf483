mkdir -p gpurun_out
cp tools/_variants/lib_gwnp.so tnmf_b200/libtnmf_b200.so
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg3 --no-cuda-graph > gpurun_out/s2j_prof.log 2>&1; echo "bench rc=$?"
grep "^gradw_ns" gpurun_out/s2j_prof.log | tail -8
