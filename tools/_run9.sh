python -m pytest tests -m gpu -x -q > gpurun_out/pytest_head3.log 2>&1; tail -3 gpurun_out/pytest_head3.log
python bench.py > gpurun_out/bench_head3.json 2> gpurun_out/bench_head3.err; tail -c 200 gpurun_out/bench_head3.json; tail -3 gpurun_out/bench_head3.err
python tools/bh_once.py 72 128; python tools/bh_once.py 72 128 1
