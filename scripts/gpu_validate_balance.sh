python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python scripts/quick_bench.py --fused 4096 8192 16384 2>&1 | tail -3
HG_FUSED_BALANCE=0 python scripts/quick_bench.py --fused 8192 16384 2>&1 | tail -2
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_grid.py -x -q -k balanced 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_grid.py -x -q -k balanced 2>&1 | grep -E "passed|failed|ERROR SUMMARY"
