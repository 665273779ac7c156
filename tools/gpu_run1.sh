mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt
(timeout 900 python tools/line_check.py 2 3 4 -- 6 7) > gpurun_out/r2_check_default.log 2>&1
(TRIXIB200_LINE_SHAPE=12 timeout 900 python tools/line_check.py 2 3 4 -- 6 7) > gpurun_out/r2_check_pp.log 2>&1
for tool in memcheck racecheck initcheck synccheck; do
  (timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/line_check.py 2) > gpurun_out/r2_san_${tool}_default.log 2>&1
  (TRIXIB200_LINE_SHAPE=12 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/line_check.py 2) > gpurun_out/r2_san_${tool}_pp.log 2>&1
done
tail -n 12 gpurun_out/r2_check_default.log gpurun_out/r2_check_pp.log
for f in gpurun_out/r2_san_*; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|shape=" $f | tail -4; done
