# Marginal cost of kernel families inside the graph-replayed DFormer-L step: drop their launches and compare ms/step.
run() { DFB200_PROFILE_SKIP=$2 python bench.py --steps 6 --warmup 3 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],2))"; }
run base ""
run no_gaa gaa_fused_fwd,gaa_fused_bwd
run no_dw7 dwconv_fwd,dwconv_bwd
run no_mlp_dw mlp_dw_fwd,mlp_dw_bwd
run no_colsum colsum
run no_ln_bwd layernorm_bwd
run no_ln layernorm_bwd,layernorm_fwd,scale_residual_layernorm_fwd
run no_pool_resize pool7_fwd,pool7_bwd,resize_fwd,resize_bwd
run no_scale_res scale_residual_fwd,scale_residual_bwd
run no_mul_act mul_fwd,mul_bwd,act_fwd,act_bwd,axpy,cast
run no_bn bn_stats,bn_apply,bn_bwd_reduce,bn_bwd_apply
run no_loss upsample_ce_fwd,upsample_ce_bwd_fused
run no_nmf_elem mu_update,mu_update_bwd,softmax_rows,softmax_rows_bwd,normalize_cols,cast2d
run no_im2col im2col3x3s2_fwd,im2col3x3s2_bwd,unpack_conv_grad
run no_adamw adamw
run no_gemm gemm
