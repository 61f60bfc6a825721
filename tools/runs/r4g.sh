#!/bin/bash
# launch lists and full captures of the round's last kernels (r2c_* under profiles/), plus the tail_tab grid-size knob
O=gpurun_out/r4g; mkdir -p $O
python tools/ncu_cfg3_frame.py > $O/plain_cfg3.log 2>&1 || exit 1
python tools/ncu_small.py > $O/plain_small.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg3.csv python tools/ncu_cfg3_frame.py > $O/ncu_l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_small.csv python tools/ncu_small.py > $O/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pyr_rowscan|k_colscan' -s 2 -c 2 -o $O/prof_pyr -f python tools/ncu_cfg3_frame.py > $O/ncu_p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cascade_tail_tab|k_group_fused|k_stage0_rows_p|k_face_prep' -c 8 -o $O/prof_small -f python tools/ncu_small.py > $O/ncu_f.log 2>&1
for w in 0 16 32 64 128; do echo "WPB=$w $(NUBOVCA_TAILTAB_WPB=$w python tools/small_frame_latency.py 2>&1 | tail -1)"; done
ls -la $O
