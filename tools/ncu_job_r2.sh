set -x
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --frames 128 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile > gpurun_out/launches_r2.out 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"keygen|sort_pass|sort_hist|leaf_scan|leaf_emit|dec_leaves|jpeg_idct|jpeg_mcu|hist_kernel|export_kernel|assemble" -c 48 -f -o /tmp/rep/prof_par_r2 python tools/prof_step.py 1000000 16 > gpurun_out/prof_par_r2.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rc_encode_lps|rc_decode_lps|dec_jpeg" -c 8 -f -o /tmp/rep/prof_ser_lps_r2 python tools/prof_step.py 1000000 16 surf 1 > gpurun_out/prof_ser_lps_r2.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dec_entropy" -c 2 -f -o /tmp/rep/prof_ser_cta_r2 python tools/prof_step.py 1000000 16 surf 0 > gpurun_out/prof_ser_cta_r2.out 2>&1
python tools/ncu_summary.py full /tmp/rep/prof_par_r2.ncu-rep:16 /tmp/rep/prof_ser_lps_r2.ncu-rep:16 /tmp/rep/prof_ser_cta_r2.ncu-rep:16 --traffic gpurun_out/ncu_traffic_r2.json > gpurun_out/ncu_full_r2_summary.txt 2> gpurun_out/ncu_summary.err
python tools/ncu_summary.py launches gpurun_out/launches_r2.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv python bench.py --frames 128 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile" > gpurun_out/launches_r2_summary.txt 2>> gpurun_out/ncu_summary.err
# source-level hot spots of the lane-per-stream decoder (stall samples per SASS line are too big to keep: top lines only)
ncu -i /tmp/rep/prof_ser_lps_r2.ncu-rep --page source --csv -k regex:rc_decode_lps 2>/dev/null | head -400 > gpurun_out/ncu_source_lps_dec_head.csv
ls -la /tmp/rep gpurun_out | head -40
tail -3 gpurun_out/ncu_summary.err
