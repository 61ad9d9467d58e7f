#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): tests, smoke, sanitizer, bench lines for every config and both arms,
# per-op table vs the reference extension, model-level parity report, ncu launch list of the bench command, per-kernel
# ncu captures, per-forward kernel table.  Everything lands in gpurun_out/; tools/collect_profiles.py copies the
# summaries into profiles/.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -1 gpurun_out/r2_sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -1 gpurun_out/r2_sanitizer_memcheck.log
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k20_ref.json 2>/dev/null
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k20.json 2> gpurun_out/r2_bench_n1_k20.err
timeout 200 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_n1_k100.json 2>/dev/null
for c in 3 4; do
  timeout 200 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_config$c.json 2>/dev/null
  timeout 200 python bench.py --impl reference --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_config${c}_ref.json 2>/dev/null
done
timeout 200 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2_n1_config5.json 2>/dev/null
timeout 200 python bench.py --config 5 --points 200000 --steps 5 --warmup 3 > gpurun_out/r2_n1_config5_200k.json 2>/dev/null
timeout 300 python tools/bench_ops.py --graph --json gpurun_out/r2_ops_vs_reference_graph.json > gpurun_out/r2_bench_ops.log 2>&1
timeout 200 python tools/parity_report.py > gpurun_out/r2_parity.log 2>&1
timeout 120 python tools/time_sa_layers.py --json gpurun_out/r2_sa_layers.json > gpurun_out/r2_sa_layers.log 2>&1
timeout 200 python tools/marginal_cost.py --json gpurun_out/r2_marginal_cost_config2.json > gpurun_out/r2_marginal_cost.log 2>&1
timeout 100 python tools/profile_train_step.py 2>&1 | tail -40 > gpurun_out/r2_train_step_profile.txt
timeout 500 ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_forward.csv python tools/one_forward.py > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:sa_fused_pipe|sa_inline' --csv --log-file gpurun_out/r2_sa_traffic.csv python tools/one_forward.py --default-options > /dev/null 2>&1
bash tools/profile_kernels.sh
ls gpurun_out/r2_* | wc -l
