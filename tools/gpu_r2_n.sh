#!/bin/bash
# Round-2 (second session) pass 7, one GPU: sizing of the fused root directory + filter on c3 now that no L2 is set aside
# (sectors per prefix 0 = two-table path / 3 / 4 / 5; two-table path with 4, 6, 8 filter bits per k-mer).
set -x
O=gpurun_out/r2n
mkdir -p $O
A="--config c3 --steps 10 --sub '' --no-cpu-baseline --no-e2e --no-probe"
run() { # name, env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --config c3 --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/$n.json 2> $O/$n.err
}
run rkf4 BFT_B200_RKF_SECTORS=4
run rkf3 BFT_B200_RKF_SECTORS=3
run rkf5 BFT_B200_RKF_SECTORS=5
run rkf0_kf6 BFT_B200_RKF_SECTORS=0
run rkf0_kf8 BFT_B200_RKF_SECTORS=0 BFT_B200_KF_BITS=8
run rkf0_kf4 BFT_B200_RKF_SECTORS=0 BFT_B200_KF_BITS=4
run rkf4_kf0 BFT_B200_RKF_SECTORS=4 BFT_B200_KF_BITS=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n/*.json")):
    try:
        d=json.loads(open(f).read()); r=d["roofline"]
        print(f.split("/")[-1], "%.2f G  %.3f ms  bucket/k %.3f  rootkf %.1f MB filter %.1f MB" % (d["value"]/1e9, r["kernel_ms"], r["bucket_accesses_per_kmer"], r["rootkf_mb"], r["filter_mb"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
