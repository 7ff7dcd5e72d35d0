#!/bin/bash
# ncu captures of round 2 (run on the GPU box from the repo root): launch list of the bench command + full captures of the
# dominant kernels of the fp32 (3xTF32) path and of the JPEG ingest.  Nothing printed under ncu is used as a bench value.
# The .ncu-rep files are reduced to CSV / text on the box and deleted (gpurun_out/ is limited to 64 MiB).
set -x
OUT=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain 0 > $OUT/bench_pre_ncu.json 2> $OUT/bench_pre_ncu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_bench_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain 0 > $OUT/ncu_launch.log 2>&1
python tools/prof_effnet.py 256 1 fp32 > $OUT/prof_pre.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tf32x3|k_dw_tile_f32" -c 49 -o $OUT/ncu_fp32_r02 -f \
    python tools/prof_effnet.py 256 1 fp32 > $OUT/ncu_fp32.log 2>&1
ncu -i $OUT/ncu_fp32_r02.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py > $OUT/ncu_fp32_r02_raw.csv
python tools/ncu_traffic.py $OUT/ncu_fp32_r02_raw.csv > $OUT/traffic_r02.json
ncu -i $OUT/ncu_fp32_r02.ncu-rep --page details --kernel-name regex:k_gemm_tf32x3 --launch-count 1 --launch-skip 4 > $OUT/ncu_tf32_b1expand_details_r02.txt 2>/dev/null
ncu -i $OUT/ncu_fp32_r02.ncu-rep --page details --kernel-name regex:k_gemm_tf32x3 --launch-count 1 --launch-skip 25 > $OUT/ncu_tf32_b9project_details_r02.txt 2>/dev/null
rm -f $OUT/ncu_fp32_r02.ncu-rep
python tools/jpeg_probe.py > $OUT/jpeg_pre.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_jh_|k_ju_|k_jpeg_" --launch-skip 98 -c 14 -o $OUT/ncu_jpeg_r02 -f \
    python tools/jpeg_probe.py > $OUT/ncu_jpeg.log 2>&1
ncu -i $OUT/ncu_jpeg_r02.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py > $OUT/ncu_jpeg_r02_raw.csv
rm -f $OUT/ncu_jpeg_r02.ncu-rep
python tools/tta_overlay_probe.py > $OUT/tta_overlay_probe_r02.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_tta_hpass|k_overlay|k_clahe_hpass|k_vpass" --launch-skip 6 -c 5 -o $OUT/ncu_tta_r02 -f \
    python tools/tta_overlay_probe.py > $OUT/ncu_tta.log 2>&1
ncu -i $OUT/ncu_tta_r02.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py > $OUT/ncu_tta_overlay_r02_raw.csv
rm -f $OUT/ncu_tta_r02.ncu-rep
ls -la $OUT | tail -20
