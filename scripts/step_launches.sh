# launch list of one bench-size build (kernel shares of a step); summary -> gpurun_out/launches_step_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_step.csv python scripts/build_once.py bench > gpurun_out/launches_step.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_step.csv > gpurun_out/launches_step_summary.txt 2>&1
