python -m pytest tests/test_rips_large_gpu.py tests/test_takens_gpu.py tests/test_rips_differential_gpu.py tests/test_coupling_gpu.py -m gpu -q -x 2>&1 | tail -2
python tools/large_bench.py 2>&1 | tail -2 | cut -c1-300
