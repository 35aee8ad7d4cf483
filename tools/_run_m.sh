mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_ or narrow_atoms or random_shapes or nan_guards or graph_replay or float32_100" > gpurun_out/s2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2m_pytest.log
tail -3 gpurun_out/s2m_pytest.log
for i in 1 2 3; do timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tall_atoms or graph_replay" 2>&1 | tail -1; done
timeout 200 python tools/determinism_check.py 2>&1 | tail -4
