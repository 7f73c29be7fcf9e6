set -x
python -m pytest tests/test_tfa_losses_gpu.py -x -q 2>&1 | tail -4
python tools/tfa_ab.py /tmp/a.npz
DIF_TFA_GENERIC_HARD=1 python tools/tfa_ab.py /tmp/b.npz
python tools/tfa_ab.py --cmp /tmp/a.npz /tmp/b.npz
python tools/tfa_once.py 1024 4 128
python tools/tfa_once.py 256 4 128
python tools/tfa_once.py 18 4 128
python tools/tfa_once.py 256 16 128
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/tfa_launches3.csv python tools/tfa_once.py 1024 4 128 2 > /dev/null 2>&1
