#!/bin/bash
# ncu --set full captures of the kernels VERDICT r1 asked evidence for (run under gpurun, one GPU):
#   group_points_kernel (config-4 shape), bqg_query_kernel (SA1 ball query), fps_cluster_kernel, fps_bucket_kernel,
#   pm_linear_kernel, sa_inline_kernel / sa_fused_pipe_kernel, group_points_grad_gather_kernel.  Writes gpurun_out/r2_<name>.ncu-rep + a raw-metrics CSV each.
set -u
mkdir -p gpurun_out
prof() {  # name, kernel regex, launches of that kernel to skip, command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c 1 -f -o gpurun_out/r2_$name "$@" > gpurun_out/r2_$name.log 2>&1
  ncu -i gpurun_out/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_$name.raw.csv 2>/dev/null
}
prof group_points 'group_points_kernel' 20 python tools/bench_ops.py --ops group --no-ref --iters 3
prof bqg_query 'bqg_query_kernel' 2 python tools/bench_ops.py --ops ball_query --no-ref --iters 3
prof fps_cluster 'fps_cluster_kernel' 2 python tools/time_fps.py --algos cluster --streams 1
prof fps_bucket 'fps_bucket_kernel' 2 python tools/time_fps.py --algos bucket --streams 1
prof pm_linear 'pm_linear_kernel' 30 python tools/one_forward.py
prof sa_fused_sa2 'sa_fused_pipe_kernel' 12 python tools/one_forward.py --default-options
prof sa_fused_sa1 'sa_inline_kernel' 3 python tools/one_forward.py --default-options
prof group_grad_gather 'group_points_grad_gather_kernel' 1 python tools/bench_ops.py --ops group --no-ref --iters 3
prof three_interpolate 'three_interpolate_kernel' 8 python tools/bench_ops.py --ops interp --no-ref --iters 3
