#!/bin/bash
# Round-2 closing check on one B200 (run through gpurun): the complete GPU parity suite, smoke(), the default bench line,
# and ncu --set full captures of the two sweep kernels of the final (512-row tile) build: a dense early iteration and a
# culled late one. Every step has its own time limit; outputs go to gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
t0=$(date +%s)
timeout 480 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/gputest_r02_final.log 2>&1
echo "pytest rc=$? at $(( $(date +%s) - t0 )) s" | tee -a gpurun_out/gputest_r02_final.log
tail -5 gpurun_out/gputest_r02_final.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke_r02_final.log 2>&1
echo "smoke rc=$? at $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/smoke_r02_final.log
timeout 330 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
echo "bench rc=$? at $(( $(date +%s) - t0 )) s"; cut -c1-400 gpurun_out/bench_r02_final.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:estep_sweep -s 4 -c 2 -f \
  -o gpurun_out/prof_estep_r02_final_dense python profiles/profile_estep.py --iters 4 > gpurun_out/ncu_dense.log 2>&1
echo "ncu dense rc=$? at $(( $(date +%s) - t0 )) s"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:estep_sweep -s 340 -c 2 -f \
  -o gpurun_out/prof_estep_r02_final_late python profiles/profile_estep.py --iters 175 --nn-init 1 > gpurun_out/ncu_late.log 2>&1
echo "ncu late rc=$? at $(( $(date +%s) - t0 )) s"
ls -la gpurun_out | tail -12
