(timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "graph_capturable" 2>&1 | tail -15)
