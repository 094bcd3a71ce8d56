python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -5
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for B in 32 4; do python bench.py --batch $B --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --dump-profile gpurun_out/prof_attn2wg_b$B.csv > gpurun_out/bench_attn2wg_b$B.json 2>gpurun_out/bench_attn2wg_b$B.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_attn2wg_b$B.json').read().strip().splitlines()[-1]);print('B',$B,d['value'],d['ms_per_step'])"; grep attn gpurun_out/prof_attn2wg_b$B.csv | head -3; done
