#!/usr/bin/env bash
mkdir -p gpurun_out
P=${1:-fin}
timeout 300 python tools/lm_attn_check.py check 2>&1 | grep -E "LM check|unimp|rror" | tail -4
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/${P}_smoke.log
timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm --tag $P --no-eager 2>&1 >/dev/null | grep "^KB"
timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm --tag $P --no-eager 2>&1 >/dev/null | grep "^KB"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lm_attn_bwd_kernel' -s 2 -c 2 -o gpurun_out/${P}_ncu_lm_bwd_c2 -f python tools/kbench_cli.py --workload C2-rec --only lm --no-eager > gpurun_out/${P}_ncu_lm_bwd_c2.log 2>&1
echo "ncu lm bwd c2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lm_attn_bwd_kernel' -s 1 -c 2 -o gpurun_out/${P}_ncu_lm_bwd_c3 -f python tools/kbench_cli.py --workload C3-multitask --only lm --no-eager > gpurun_out/${P}_ncu_lm_bwd_c3.log 2>&1
echo "ncu lm bwd c3 rc=$?"
