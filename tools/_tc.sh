(cd tools/microbench && timeout 60 ./tmem_read 2>&1 | grep -v "no fold")
for c in 0 1 2 3 4; do
URED_TC_CONFIG=$c timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config $c: cfg2 ms', round(d['ms_per_step'],4), 'nn ms', round(d['roofline']['kernel_ms'],4))"
done
URED_TC_CONFIG=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_op.py -q -m gpu 2>&1 | tail -2
URED_TC_CONFIG=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_op.py -q -m gpu 2>&1 | tail -2
