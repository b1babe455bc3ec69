#!/bin/bash
# Rebuilds the library on the GPU box with different k_inflate table sizes / occupancy targets and times the BGZF path.
# usage (under gpurun): bash profiles/probes/inflate_variants.sh "LIT DIST MINBLOCKS" ...
cd "$(dirname "$0")/../../bamqc_b200/csrc" || exit 1
for cfg in "$@"; do
  set -- $cfg
  touch engine.cu
  make -j16 EXTRA="-DBQC_INFLATE_LITBITS=$1 -DBQC_INFLATE_DISTBITS=$2 -DBQC_INFLATE_MINBLOCKS=$3" > /tmp/mk.log 2>&1 || { tail -5 /tmp/mk.log; continue; }
  for st in 1 2; do
    (cd ../.. && BQC_INFLATE_STREAMS=$st python bench.py --steps 3 --warmup 3 --no-cpu-baseline --bgzf-records ${BGZF_RECORDS:-3300000} > gpurun_out/infv_$1_$2_$3_s$st.json 2> gpurun_out/infv_$1_$2_$3_s$st.err; echo "lit=$1 dist=$2 minblocks=$3 streams=$st: $(grep -h e2e_bgzf gpurun_out/infv_$1_$2_$3_s$st.err)")
  done
done
