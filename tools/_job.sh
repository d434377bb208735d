timeout 300 python tools/p2_probe.py 2>&1 | grep -E "wall|p2_resolve|^ +[0-9]"
timeout 900 python -m pytest tests/test_gpu_recommend.py -m gpu -x -q 2>&1 | tail -2
