#!/bin/bash
# final checks on the final build (no bench data needed): the whole -m gpu suite, then compute-sanitizer memcheck over the tests that
# drive the new kernels (fused wide rows, collapsed subtrees incl. tight blocks, fused root directory + filter, REDUX staging /
# register counters of the sequence kernel)
O=gpurun_out/r2t
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py -x -q -m gpu -k "golden_kmers or golden_sequences or golden_branching or device_resident or (accelerator_tables and (DEEP_TIGHT or NO_DEEP))" > $O/sanitizer_memcheck_golden.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck_golden.log
tail -6 $O/sanitizer_memcheck_golden.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kmer_presence_and_colours" > $O/sanitizer_memcheck_wide_rows.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck_wide_rows.log
tail -6 $O/sanitizer_memcheck_wide_rows.log
