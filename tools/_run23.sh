python -m pytest tests/test_tfa_losses_gpu.py tests/test_losses_gpu.py -x -q 2>&1 | tail -2
python tools/tfa_ab.py /tmp/a.npz
DIF_TFA_SLOW_PDIST=1 python tools/tfa_ab.py /tmp/b.npz
python tools/tfa_ab.py --cmp /tmp/a.npz /tmp/b.npz
python tools/tfa_step_once.py 1024 4 128
python tools/tfa_step_once.py 1024 4 512
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:canon_mm -s 1 -c 2 python tools/tfa_once.py 1024 4 128 2 2>&1 | grep -E "gpu__time"
