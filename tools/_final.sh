set -x
python tools/bench_kernels.py > gpurun_out/kernels_c3_final.jsonl 2>&1
python tools/bench_kernels.py --d 384 > gpurun_out/kernels_c3_d384_final.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_grad_pair -c 1 -f -o gpurun_out/prof_onepass_c3 python tools/bench_kernels.py --only tc_score_onepass > gpurun_out/ncu_onepass.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mask_topk_kernel -c 1 -f -o gpurun_out/prof_mask_topk_final python tools/bench_topk_fp32.py > gpurun_out/ncu_topk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_ln_fwd -c 1 -f -o gpurun_out/prof_gather_final python tools/bench_gather.py > gpurun_out/ncu_gather.log 2>&1
python tools/bench_topk_fp32.py > gpurun_out/topk_fp32_final.jsonl 2>&1
python tools/bench_gather.py > gpurun_out/gather_final.jsonl 2>&1; python tools/bench_gather.py --uniform >> gpurun_out/gather_final.jsonl 2>&1
