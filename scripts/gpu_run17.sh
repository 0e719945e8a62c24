mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -k "heat_3d or heat_2d or heat_1d or elasticity or manufactured or smoke" 2>&1 | tail -4 ) > gpurun_out/r02_t17.log 2>&1; cat gpurun_out/r02_t17.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu > gpurun_out/r02_bench_p.json 2>gpurun_out/r02_bench_p.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_p.json').read()); print(d['ms_per_step'], d['cg_iters_per_step'], d['roofline_step']); e=d['elasticity']; print(e['solve_ms'], e['cg_iters'], e['roofline']['frac'])"
