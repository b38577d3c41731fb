#!/bin/bash
# compute-sanitizer passes over a small end-to-end solve and the kernel-level tests (run on a GPU box; not part of pytest):
#     gpurun --timeout 900 -- 'bash tests/gpu_sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the TMA ring, the pair-forming tiles);
# synccheck: barrier misuse; initcheck: reads of uninitialised device memory.  Kernels run 10-100x slower under the tools,
# so the inputs are small; the multi-GPU exchange kernels spin on peer flags and are not run under the tools.
set -u
cd "$(dirname "$0")/.."
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='fused_update_apply_Hv or apply_Hv_gram or level1 or objective_and_fused_trial'
for tool in memcheck racecheck synccheck initcheck; do
    echo "=== $tool ==="
    timeout 600 $SAN --tool $tool --error-exitcode 9 --print-limit 20 \
        python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL and not 1048" 2>&1 | tail -15
    echo "exit: $?"
done
echo "=== memcheck: whole solves (host-driven, resident graph, L-BFGS-B) ==="
timeout 600 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python tests/quick_sanitize_target.py 2>&1 | tail -15
