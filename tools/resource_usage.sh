#!/bin/bash
# Registers / spills (STACK) / shared memory per kernel of libdformer_b200.so, hot kernels first (no GPU needed):
#   tools/resource_usage.sh > profiles/r01_kernel_resources.txt
LIB=${1:-dformer_b200/libdformer_b200.so}
echo "cuobjdump --dump-resource-usage $LIB  (sm_100a)"
printf "%5s %6s %8s %8s  %s\n" REG STACK SHARED CONST0 kernel
cuobjdump --dump-resource-usage "$LIB" 2>/dev/null | awk '
  /Function/ { fn=$2; sub(/:$/, "", fn); next }
  /REG:/ { r=s=sh=c=0
    for (i=1;i<=NF;i++) { split($i,a,":"); if (a[1]=="REG") r=a[2]; if (a[1]=="STACK") s=a[2]; if (a[1]=="SHARED") sh=a[2]; if (a[1]=="CONSTANT[0]") c=a[2] }
    printf "%5d %6d %8d %8d  %s\n", r, s, sh, c, fn }' | c++filt | grep -E "gemm_tc|mlp_dw|dw7_|gaa_fused|ln_(fwd|bwd)_vec|scale_residual|upsample_ce|resize_bwd_block|adamw|train_pre|mu_update|bn_" | cut -c1-200 | sort -k5 | uniq
