#!/bin/bash
# A/B timing of the all-pairs pipeline shapes (env knobs of plan_run); prints value / kernel ms per variant.
set -u
mkdir -p gpurun_out
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --workload dist --no-cpu-baseline --no-extra --steps 5 --warmup 3 2> gpurun_out/ab_$label.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$label', 'value %.4g' % d['value'], 'kernel_ms %.2f' % d['details']['step_breakdown_ms']['all_pairs_kernel'], 'e2e %.4g' % d['e2e']['value'])" || tail -5 gpurun_out/ab_$label.err
}
jm() { local label=$1; shift; env "$@" timeout 600 python scripts/jmle_run.py ${JP:-16} ${JN:-4000} 3 2>&1 | tail -1 | sed "s/^/$label /"; }
for v in "$@"; do
  case $v in
    u3) run u3_s7 DB200_SWEEP_CTAS=3 DB200_SWEEP_STAGES=7;;
    u3s6) run u3_s6 DB200_SWEEP_CTAS=3 DB200_SWEEP_STAGES=6;;
    u3s8) run u3_s8 DB200_SWEEP_CTAS=3 DB200_SWEEP_STAGES=8;;
    u2) run u2_s8 DB200_SWEEP_CTAS=2 DB200_SWEEP_STAGES=8;;
    u2s12) run u2_s12 DB200_SWEEP_CTAS=2 DB200_SWEEP_STAGES=12;;
    u3one) run u3_onechunk DB200_SWEEP_CTAS=3 DB200_SCRATCH_MB=8192;;
    u3small) run u3_128mb DB200_SWEEP_CTAS=3 DB200_SCRATCH_MB=128;;
    j2) jm j2 DB200_SWEEP_CTAS=2 DB200_SWEEP_STAGES=8;;
    j3) jm j3 DB200_SWEEP_CTAS=3 DB200_SWEEP_STAGES=6;;
    j2p14) JP=14 JN=8000 jm j2p14 DB200_SWEEP_CTAS=2;;
  esac
done
