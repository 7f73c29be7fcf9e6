set -x
python -m pytest tests/test_losses_gpu.py -x -q 2>&1 | tail -3
python tools/bh_once.py 4096 128
python tools/bh_once.py 4096 128 1
python tools/bh_once.py 1024 128
python tools/bh_once.py 8192 128
