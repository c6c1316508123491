set -x
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attn_fwd_sm100 -s 2 -c 1 -o gpurun_out/r2b_fwd40 python scripts/ncu_fwd.py 2 > gpurun_out/r2b_ncu_fwd40.log 2>&1
timeout 300 $NCU -k regex:attn_fwd_sm100 -s 2 -c 1 -o gpurun_out/r2b_fwd80 python scripts/ncu_fwd.py 2 3 1024 80 > gpurun_out/r2b_ncu_fwd80.log 2>&1
timeout 300 $NCU -k regex:attn_bwd64 -s 2 -c 1 -o gpurun_out/r2b_bwd40 python scripts/ncu_bwd.py > gpurun_out/r2b_ncu_bwd40.log 2>&1
timeout 300 $NCU -k regex:removal_corr -s 1 -c 1 -o gpurun_out/r2b_corr410 python scripts/ncu_corr.py > gpurun_out/r2b_ncu_corr410.log 2>&1
timeout 300 $NCU -k regex:removal_corr -s 1 -c 1 -o gpurun_out/r2b_corr76 python scripts/ncu_corr.py 8 4096 40 76 > gpurun_out/r2b_ncu_corr76.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv python scripts/profile_edit.py --steps 10 > gpurun_out/r2b_ncu_launches.log 2>&1
timeout 60 scripts/_bin/fwd_trace 128 > gpurun_out/r2b_trace128_final.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
