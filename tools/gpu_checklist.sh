#!/usr/bin/env bash
# First GPU call of a round: everything that has to be (re)measured on a B200, in the order of
# what blocks what.  Run from the repository root on the GPU box, e.g.
#   gpurun --timeout 900 -- 'bash tools/gpu_checklist.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
step() { echo "== $*" | tee -a gpurun_out/checklist.log; }

step "1. tests that have never run on a B200 (reference-form goldens, spherical degree 4)"
timeout 120 python -m pytest tests/test_zz_gpu_reference_form.py -m gpu -q -rA 2>&1 | tail -25 | tee gpurun_out/zz_reference_form.log

step "2. the whole GPU suite"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/gpu_suite.log

step "3. smoke + bench line (C1, N=1)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json

step "4. beyond-L2 roofline numbers (C2: 5 M cells, 12 species)"
timeout 400 python tools/measure_large.py --n 94 > gpurun_out/c2_n94.json 2> gpurun_out/c2_n94.err; tail -c 800 gpurun_out/c2_n94.json

step "5. launch list of the bench command (shares per kernel), only after step 3 exited 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
echo done | tee -a gpurun_out/checklist.log
