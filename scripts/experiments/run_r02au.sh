python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 --groups 1 > gpurun_out/r02au_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02au_launches.csv python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 --groups 1 > gpurun_out/r02au_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02au_launches.csv 2>/dev/null | head -6
