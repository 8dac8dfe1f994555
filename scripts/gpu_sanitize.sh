#!/bin/bash
# compute-sanitizer over a small-size subset of the GPU parity tests (memcheck + racecheck)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
SEL='test_msm_dma_vs_oracle and (33 or 1000) or test_msm_merged_table_vs_oracle and BN254 or test_msm_window_sizes or test_msm_merged_table_tiny_inputs or test_msm_hbm_mode_and_labels'
(time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_msm_gpu.py -m gpu -x -q -k "$SEL") > $O/sanitize_memcheck_msm.log 2>&1
echo "memcheck msm rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitize_memcheck_msm.log | tail -3
# round 2: multi-device client, batched-affine sweep (BN254: the smallest field), arena, register file, published values
SEL2='test_group_dma_vs_oracle and BN254 and (5 or 1000) or test_group_hbm_chunked or test_batched_affine_sweep_vs_oracle and BN254 and 3 or test_dma_input_replaced or test_get_api or test_external_kats_through_cuda or test_ntt_rejects'
(time timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_round2_gpu.py -m gpu -x -q -k "$SEL2") > $O/sanitize_memcheck_round2.log 2>&1
echo "memcheck round2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitize_memcheck_round2.log | tail -3
(time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_ntt_gpu.py tests/test_poseidon_gpu.py -m gpu -x -q -k "vs_oracle or edge or double_buffer or height_5 or published or random_tree") > $O/sanitize_memcheck_ntt.log 2>&1
echo "memcheck ntt+poseidon rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitize_memcheck_ntt.log | tail -3
(time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_msm_gpu.py tests/test_poseidon_gpu.py -m gpu -x -q -k "test_msm_window_sizes or test_msm_merged_table_tiny_inputs or published") > $O/sanitize_racecheck_msm.log 2>&1
echo "racecheck msm rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/sanitize_racecheck_msm.log | tail -3
