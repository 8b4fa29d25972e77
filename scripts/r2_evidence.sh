#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): the contract bench line, the ncu launch list of one omni step, `--set full`
# captures of the attention kernels (tower, fusion-encoder cross- and self-attention) and of the tower GEMMs at the omni
# step's 197 376-token shapes.  The reports are summarised ON the box (ncu -i needs no GPU) and only the text summaries are
# kept: gpurun_out/ is capped at 64 MiB.
T=${1:-r2i}
O=gpurun_out
python __graft_entry__.py --smoke > $O/${T}_smoke.txt 2>&1; tail -2 $O/${T}_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 --gpu-eager > $O/${T}_bench_omni.json 2> $O/${T}_bench_omni.err
# launch list: two plain steps, the second one captured (cudaProfilerStart / Stop around it)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file /tmp/${T}_launches.csv \
    python scripts/omni_shapes.py --plain 2 > $O/${T}_ncu_launches.log 2>&1
python scripts/launch_summary.py /tmp/${T}_launches.csv "omni step (bench.py defaults: bs 64, video 8 + audio 3 + depth 1 + text 128), second step of scripts/omni_shapes.py --plain 2 under ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off" > $O/${T}_launches_summary.txt 2>> $O/${T}_ncu_launches.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_ -c 20 -f -o /tmp/${T}_attn \
    python scripts/profile_kernels.py attn2 > $O/${T}_ncu_attn.log 2>&1
python scripts/ncu_summary.py /tmp/${T}_attn.ncu-rep > $O/${T}_ncu_attention_summary.txt 2>> $O/${T}_ncu_attn.log
for i in 6 8 10 12; do python scripts/ncu_hot.py /tmp/${T}_attn.ncu-rep $i 18 >> $O/${T}_ncu_attention_hot.txt 2>> $O/${T}_ncu_attn.log; done
MICO_PROFILE_FRAMES=768 timeout 600 ncu --set full --clock-control none -k regex:gemm_bf16_kernel -s 8 -c 13 -f -o /tmp/${T}_gemm \
    python scripts/profile_kernels.py gemm > $O/${T}_ncu_gemm.log 2>&1
python scripts/ncu_summary.py /tmp/${T}_gemm.ncu-rep > $O/${T}_ncu_gemm_summary.txt 2>> $O/${T}_ncu_gemm.log
python scripts/ncu_to_traffic.py /tmp/${T}_gemm.ncu-rep $O/${T}_gemm_traffic.json "tower GEMMs at the omni step's 197 376-token pass: fc1 fwd (+GELU, GELU' store), fc2 dgrad (x GELU'), fc2 wgrad, fc2 fwd (+residual); then proj fwd, qkv wgrad, fc1 fwd (activation only), three times each" 4 >> $O/${T}_ncu_gemm.log 2>&1
ls -la $O /tmp/${T}_*; du -sh $O
