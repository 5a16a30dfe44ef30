#!/bin/bash
# run 31: ncu --set full of the kernels that changed after run 19 (persistent attention backward, one-pass LoKr final, GroupNorm)
O=gpurun_out/run31; mkdir -p $O
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel --launch-skip 1 --launch-count 1 -o $O/attn_bwd -f python tools/profile_one.py attn > $O/ncu_attn_bwd.log 2>&1; tail -1 $O/ncu_attn_bwd.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lokr_fused --launch-skip 1 --launch-count 1 -o $O/lokr_fused -f python tools/profile_one.py lokr_fused > $O/ncu_lokr.log 2>&1; tail -1 $O/ncu_lokr.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"ln_bwd_stream|gn_fwd_fused|gn_bwd_fused|ln_fwd" --launch-count 8 -o $O/norms -f python tools/profile_one.py bw > $O/ncu_norms.log 2>&1; tail -1 $O/ncu_norms.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 1 --launch-count 1 -o $O/gemm_res -f python tools/profile_one.py gemm_res > $O/ncu_gemm_res.log 2>&1; tail -1 $O/ncu_gemm_res.log
echo DONE
