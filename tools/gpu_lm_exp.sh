#!/usr/bin/env bash
timeout 300 python tools/lm_attn_check.py check 2>&1 | grep -E "LM check|unimp|rror" | tail -3
for f in 0 4; do
  echo "UNIMP_LM_BWD_FLAGS=$f"
  UNIMP_LM_BWD_FLAGS=$f timeout 300 python tools/kbench_cli.py --workload C3-multitask --only lm --tag f$f --no-eager 2>&1 >/dev/null | grep "^KB.*bwd"
  UNIMP_LM_BWD_FLAGS=$f timeout 300 python tools/kbench_cli.py --workload C2-rec --only lm --tag f$f --no-eager 2>&1 >/dev/null | grep "^KB.*bwd"
done
