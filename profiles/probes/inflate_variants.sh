#!/bin/bash
# Rebuilds the library on the GPU box with different k_inflate table sizes / occupancy targets and times the BGZF path.
# usage (under gpurun): bash profiles/probes/inflate_variants.sh "LITBITS DISTBITS WARPS_PER_CTA MIN_CTAS_PER_SM" ...
cd "$(dirname "$0")/../../bamqc_b200/csrc" || exit 1
for cfg in "$@"; do
  set -- $cfg
  touch engine.cu
  make -j16 EXTRA="-DBQC_INFLATE_LITBITS=$1 -DBQC_INFLATE_DISTBITS=$2 -DBQC_INFLATE_WARPS=$3 -DBQC_INFLATE_MINBLOCKS=$4" > /tmp/mk.log 2>&1 || { tail -5 /tmp/mk.log; continue; }
  for st in ${STREAMS:-1 2}; do
    (cd ../.. && BQC_INFLATE_STREAMS=$st python bench.py --steps 3 --warmup 3 --no-cpu-baseline --bgzf-records ${BGZF_RECORDS:-3300000} > gpurun_out/infv_$1_$2_$3_$4_s$st.json 2> gpurun_out/infv_$1_$2_$3_$4_s$st.err; echo "lit=$1 dist=$2 warps=$3 minblocks=$4 streams=$st: $(grep -h e2e_bgzf gpurun_out/infv_$1_$2_$3_$4_s$st.err)")
  done
done
