python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_size" 2>&1 | tail -3
echo "=== 2CTA gemm tests"
NRW_TC_2CTA=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm" 2>&1 | tail -5
echo "=== 2CTA parity"
NRW_TC_2CTA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tcgen05 or golden or networks" 2>&1 | tail -3
for c in 131072 262144; do
python bench.py --steps 5 --chunk_rows $c --no_cpu_baseline --no_torch_gpu_ref 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1CTA chunk', $c, 'ms', round(d['ms_per_step'],1), 'rays/s', round(d['value']), 'launches', d['gpu_launches'])"
NRW_TC_2CTA=1 timeout 300 python bench.py --steps 5 --chunk_rows $c --no_cpu_baseline --no_torch_gpu_ref 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2CTA chunk', $c, 'ms', round(d['ms_per_step'],1), 'rays/s', round(d['value']), 'launches', d['gpu_launches'])"
done
