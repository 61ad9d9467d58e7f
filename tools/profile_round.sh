set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sa_fused_pipe --launch-skip 10 --launch-count 5 -f -o gpurun_out/prof_sa_r1c python tools/one_forward.py > gpurun_out/ncu_sa_r1c.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r1c.log 2>&1
python bench.py --impl reference > gpurun_out/bench_r1c_ref.json 2> gpurun_out/bench_r1c_ref.err
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
tail -c 600 gpurun_out/bench_r1c_ref.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1c.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['clocks'], d['hbm_ops'], d['roofline']['frac'])"
