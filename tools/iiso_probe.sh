run() { timeout 300 python bench.py "$@" --steps 100 --warmup 10 --no-variants --no-e2e --no-cpu-baseline --no-like-for-like --c4 off | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), d['config']['kernel'], round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms_per_launch']*1e3,1), 'us')"; }
for i in 1 2; do
echo -n "f64 iiso dif2: "; run --dtype f64 --update-type 3 --dif-order 2
echo -n "f32 iiso dif2: "; run --dtype f32 --update-type 3 --dif-order 2
echo -n "f32 iiso dif0: "; run --dtype f32 --update-type 3 --dif-order 0
echo -n "f64 iiso dif0: "; run --dtype f64 --update-type 3 --dif-order 0
done
python -m pytest tests/test_gpu_interp.py tests/test_gpu_dif.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
