set -x
python tools/copy_calib.py > gpurun_out/r2j8_copy.log 2>&1; cat gpurun_out/r2j8_copy.log
