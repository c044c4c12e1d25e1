#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/debug_nondet.py 4 64 12 > gpurun_out/r18_nondet.log 2>&1; cat gpurun_out/r18_nondet.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_boundary.py::test_two_handles_step_concurrently_and_bit_exactly > gpurun_out/r18_pytest.log 2>&1; tail -8 gpurun_out/r18_pytest.log | cut -c1-300
