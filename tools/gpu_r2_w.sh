#!/bin/bash
# compute-sanitizer memcheck on the final build (paired buckets): golden look-ups through every entry point, wide rows, sequences
O=gpurun_out/r2w
mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py -x -q -m gpu -k "golden_kmers or golden_sequences or golden_branching or device_resident or reference_record_format or (accelerator_tables and (DEEP_TIGHT or NO_DEEP))" > $O/sanitizer_memcheck_golden.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck_golden.log
tail -4 $O/sanitizer_memcheck_golden.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kmer_presence_and_colours or enumeration" > $O/sanitizer_memcheck_parity.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck_parity.log
tail -4 $O/sanitizer_memcheck_parity.log
